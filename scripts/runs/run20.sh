set -u
O=gpurun_out/r3a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
# 1. the new schedule / kernels against their references
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or bilinear" > $O/pytest_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -3 $O/pytest_gemm.log | cut -c1-400
timeout 600 python -m pytest tests/test_glue_kernels_gpu.py -m gpu -q > $O/pytest_glue.log 2>&1; echo "glue tests rc=$?"; tail -3 $O/pytest_glue.log | cut -c1-400
timeout 900 python -m pytest tests/test_big_goldens_gpu.py -m gpu -q -k "cfg4_b4 or ip_cfg3" > $O/pytest_big.log 2>&1; echo "big goldens rc=$?"; tail -3 $O/pytest_big.log | cut -c1-400
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -k "training_step or graph_replay" > $O/pytest_train.log 2>&1; echo "train tests rc=$?"; tail -3 $O/pytest_train.log | cut -c1-400
# 2. A/B of the stream-K schedule on this box (same clocks / power cap)
for sk in 0 1 0 1; do
  MTT_GEMM_STREAMK=$sk timeout 300 python bench.py --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-train-leg > $O/ab_sk${sk}_$RANDOM.json 2>> $O/ab.err; echo "ab sk=$sk rc=$?"
done
MTT_GEMM_STREAMK=1 timeout 300 python bench.py --config ip_cfg3 --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline > $O/ip_sk1.json 2>> $O/ab.err; echo "ip sk1 rc=$?"
# 3. the default line with the train_step block, and the training line with / without stream-K
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -2 $O/bench_default.err | cut -c1-300
MTT_GEMM_STREAMK=1 timeout 600 python bench.py --train --steps 10 --warmup 3 --repeats 3 --no-gpu-eager > $O/train_sk1.json 2> $O/train_sk1.err; echo "train sk1 rc=$?"
python - <<P
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f.split("/")[-1], round(d.get("value"),1), round(d.get("ms_per_step"),3), (d.get("clocks") or {}).get("sm_mhz"), round(r.get("frac") or 0,4), (r.get("backbone_gemms") or {}).get("avg_launch_us"), (d.get("train_step") or {}))
    except Exception as e: print(f, "FAIL", e)
P
