set -u
O=gpurun_out/r2o; mkdir -p $O
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_swin_gpu.py -m gpu -q -x > $O/pytest_train.log 2>&1; echo "train rc=$?"; grep -E "passed|failed|^E" $O/pytest_train.log | cut -c1-1200 | tail -12
timeout 1500 python bench.py --train --config tp_cfg4 --batch 4 --steps 5 --warmup 3 --repeats 1 --no-gpu-eager > $O/train_cfg4.json 2> $O/train_cfg4.err; echo "bench cfg4 rc=$?"; tail -3 $O/train_cfg4.err | cut -c1-600
python - <<P
import json
try:
    d=json.loads(open("$O/train_cfg4.json").read().strip().splitlines()[-1]); print({k:d[k] for k in ("value","ms_per_step","phases","launches_per_step")}); print({k:d["roofline"][k] for k in ("achieved","frac","tensor_ms_per_step","share_of_step")})
    for t in d["top_tensor_shapes"]: print(t)
except Exception as e: print("FAIL", e)
P
