set -u
O=gpurun_out/r2p; mkdir -p $O
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -x > $O/pytest_train.log 2>&1; echo "train rc=$?"; grep -E "passed|failed|^E" $O/pytest_train.log | cut -c1-1200 | tail -12
for flag in "--no-graph" ""; do
timeout 1500 python bench.py --train --config tp_cfg4 --batch 4 --steps 5 --warmup 3 --repeats 1 --no-gpu-eager $flag > $O/train_cfg4$flag.json 2> $O/train_cfg4$flag.err; echo "bench cfg4 $flag rc=$?"; tail -3 $O/train_cfg4$flag.err | cut -c1-600
python - <<P
import json
try:
    d=json.loads(open("$O/train_cfg4$flag.json").read().strip().splitlines()[-1]); print("$flag", {k:d[k] for k in ("value","ms_per_step","phases","launches_per_step")}); print({k:d["roofline"][k] for k in ("achieved","frac","tensor_ms_per_step","share_of_step")})
except Exception as e: print("FAIL", e)
P
done
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/ncu_train.py tp_cfg4 4 > $O/ncu_train.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_train.log
python scripts/summarize_ncu.py --launches $O/train_launches.csv $O/train_launch_shares.md; head -32 $O/train_launch_shares.md
