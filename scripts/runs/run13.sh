set -u
O=gpurun_out/r2m; mkdir -p $O
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -k "full_width or torch_facing" > $O/pytest_train.log 2>&1; echo "train rc=$?"; grep -E "passed|failed|^E" $O/pytest_train.log | cut -c1-1200 | tail -12
timeout 900 python bench.py --train --config tp_cfg4_d4 --batch 2 --steps 5 --warmup 3 --repeats 1 > $O/train_cfg4d4.json 2> $O/train_cfg4d4.err; echo "bench d4 rc=$?"; tail -3 $O/train_cfg4d4.err | cut -c1-600
python - <<P
import json
try:
    d=json.loads(open("$O/train_cfg4d4.json").read().strip().splitlines()[-1]); print({k:d[k] for k in ("value","ms_per_step","phases","launches_per_step")}); print(d["roofline"]); print(d.get("gpu_eager_baseline"))
except Exception as e: print("FAIL", e)
P
timeout 1500 python bench.py --train --config tp_cfg4 --batch 4 --steps 5 --warmup 3 --repeats 1 > $O/train_cfg4.json 2> $O/train_cfg4.err; echo "bench cfg4 rc=$?"; tail -3 $O/train_cfg4.err | cut -c1-600
python - <<P
import json
try:
    d=json.loads(open("$O/train_cfg4.json").read().strip().splitlines()[-1]); print({k:d[k] for k in ("value","ms_per_step","phases","launches_per_step")}); print(d["roofline"]); print(d.get("gpu_eager_baseline"))
except Exception as e: print("FAIL", e)
P
