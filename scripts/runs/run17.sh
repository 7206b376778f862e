set -u
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x > $O/pytest_train.log 2>&1; echo "train rc=$?"; grep -E "passed|failed|^E" $O/pytest_train.log | cut -c1-1200 | tail -12
timeout 900 python bench.py --train --config tp_cfg4 --batch 4 --steps 10 --warmup 3 --repeats 3 --no-gpu-eager > $O/train_cfg4_n1.json 2> $O/train_cfg4_n1.err; echo "n1 rc=$?"; tail -2 $O/train_cfg4_n1.err | cut -c1-400
for flag in "" "--no-graph"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --train --gpus 2 --config tp_cfg4 --batch 4 --steps 10 --warmup 3 --repeats 3 --no-gpu-eager $flag > $O/train_cfg4_n2$flag.json 2> $O/train_cfg4_n2$flag.err; echo "n2 $flag rc=$?"; grep -v "^\*\|OMP_NUM" $O/train_cfg4_n2$flag.err | tail -4 | cut -c1-500
done
python - <<P
import json,glob
for f in sorted(glob.glob("$O/train_cfg4_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d[k] for k in ("value","ms_per_step","n_gpus","phases")}, d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "FAIL", e)
P
