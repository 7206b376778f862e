set -u
O=gpurun_out/r3h; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log | cut -c1-400
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --train --steps 10 --warmup 3 --repeats 3 > $O/train_cfg4.json 2> $O/train_cfg4.err; echo "train rc=$?"; tail -2 $O/train_cfg4.err | cut -c1-300
timeout 300 python bench.py --config ip_cfg3 --steps 20 --warmup 5 --no-cpu-baseline > $O/ip_cfg3.json 2>> $O/other.err; echo "ip rc=$?"
timeout 300 python bench.py --config tp_cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --no-train-leg > $O/tp_cfg2.json 2>> $O/other.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --config tp_cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/tp_cfg5.json 2>> $O/other.err; echo "cfg5 rc=$?"
timeout 300 python bench.py --config tps_swinB --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/tps_swinB.json 2>> $O/other.err; echo "swinB rc=$?"
timeout 300 python bench.py --batch 1 --steps 40 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-train-leg > $O/cfg4_bs1.json 2>> $O/other.err; echo "bs1 rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg4.csv python scripts/ncu_forward.py tp_cfg4 parity 4 > $O/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python scripts/ncu_train.py tp_cfg4 4 > $O/ncu_train.log 2>&1; echo "ncu train rc=$?"
python - <<P
import json
for f in ("bench_default","bench_reference","train_cfg4","ip_cfg3","tp_cfg2","tp_cfg5","tps_swinB","cfg4_bs1"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d.get("value"),2), round(d.get("ms_per_step"),3), round((d.get("e2e") or {}).get("value") or 0,1), (d.get("clocks") or {}).get("sm_mhz"), round((d.get("roofline") or {}).get("frac") or 0,4), {k:round(v,1) for k,v in (d.get("gpu_eager_baseline") or {}).items() if k in ("fp32","tf32","bf16_autocast")}, (d.get("cpu_baseline") or {}).get("value"), (d.get("train_step") or {}).get("value"))
    except Exception as e: print(f, "FAIL", e)
P
cp gpurun_out/parity_r2.json $O/ 2>/dev/null
