set -u
O=gpurun_out/r3b; mkdir -p $O
timeout 300 python scripts/sk_microbench.py > $O/sk_microbench.txt 2>&1; echo "microbench rc=$?"; cat $O/sk_microbench.txt | tail -14
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "streamk or bilinear" > $O/pytest_k.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/pytest_k.log | cut -c1-300
for sk in 0 1 2; do
  MTT_GEMM_STREAMK=$sk timeout 600 python bench.py --train --steps 10 --warmup 3 --repeats 3 --no-gpu-eager > $O/train_sk$sk.json 2> $O/train_sk$sk.err; echo "train sk$sk rc=$?"
done
for sk in 0 1; do
  MTT_GEMM_STREAMK=$sk timeout 300 python bench.py --batch 1 --steps 40 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-train-leg > $O/bs1_sk$sk.json 2>> $O/ab.err; echo "bs1 sk=$sk rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-train-leg > $O/cfg4.json 2>> $O/ab.err; echo "cfg4 rc=$?"
timeout 300 python bench.py --config ip_cfg3 --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline > $O/ip_cfg3.json 2>> $O/ab.err; echo "ip rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg4.csv python scripts/ncu_forward.py tp_cfg4 parity 4 > $O/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ip_cfg3.csv python scripts/ncu_forward.py ip_cfg3 parity 4 > $O/ncu_ip.log 2>&1; echo "ncu ip rc=$?"
python - <<P
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("/")[-1], round(d.get("value"),1), round(d.get("ms_per_step"),3), (d.get("clocks") or {}).get("sm_mhz"), round(r.get("frac") or 0,4), (r.get("backbone_gemms") or {}).get("avg_launch_us"))
    except Exception as e: print(f, "FAIL", e)
P
