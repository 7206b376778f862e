set -u
O=gpurun_out/r2r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --train --config tp_cfg4 --batch 4 --steps 10 --warmup 3 --repeats 3 > $O/train_cfg4.json 2> $O/train_cfg4.err; echo "train rc=$?"; tail -2 $O/train_cfg4.err | cut -c1-300
timeout 600 python bench.py --train --mode speed --config tp_cfg4 --batch 4 --steps 10 --warmup 3 --repeats 3 --no-gpu-eager > $O/train_cfg4_speed.json 2> $O/train_cfg4_speed.err; echo "train speed rc=$?"; tail -2 $O/train_cfg4_speed.err | cut -c1-300
python - <<P
import json
for f in ("bench_default","bench_reference","train_cfg4","train_cfg4_speed"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("roofline") or {}).get("frac"), {k:v for k,v in (d.get("gpu_eager_baseline") or {}).items() if k in ("fp32","tf32","bf16_autocast")}, (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "FAIL", e)
P
cp gpurun_out/parity_r2.json $O/ 2>/dev/null
