set -u
O=gpurun_out/r2s; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -k "graph_replay" > $O/pytest_graph.log 2>&1; echo "graph test rc=$?"; tail -2 $O/pytest_graph.log | cut -c1-300
timeout 600 python bench.py --train --config tp_cfg2 --batch 4 --steps 10 --warmup 3 --repeats 3 > $O/train_cfg2.json 2> $O/train_cfg2.err; echo "train cfg2 rc=$?"; tail -2 $O/train_cfg2.err | cut -c1-300
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"attn_softmax_bwd|transpose_planes|transpose_split|im2col3x3_t|adam_kernel" -c 12 -o $O/train_glue python scripts/ncu_train.py tp_cfg4 4 > $O/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_full.log
python - <<P
import json
try:
    d=json.loads(open("$O/train_cfg2.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["frac"], {k:v for k,v in d["gpu_eager_baseline"].items() if k in ("fp32","tf32","bf16_autocast")})
except Exception as e: print("FAIL", e)
P
