set -u
O=gpurun_out/r3f; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -3 $O/bench_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<P
import json
for f in ("bench_n2","bench_ref_n2"):
    try:
        d=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("n_gpus"), (d.get("e2e") or {}).get("value"), d.get("train_step"))
    except Exception as e: print(f, "FAIL", e)
P
