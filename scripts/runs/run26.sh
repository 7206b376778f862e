set -u
O=gpurun_out/r3g; mkdir -p $O
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q > $O/pytest_train.log 2>&1; echo "train tests rc=$?"; tail -3 $O/pytest_train.log | cut -c1-400
timeout 600 python bench.py --train --steps 10 --warmup 3 --repeats 3 --no-gpu-eager > $O/train_cfg4.json 2> $O/train_cfg4.err; echo "train rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"im2col3x3_t|transpose|split_kernel|colreduce" --log-file $O/train_glue_launches.csv python scripts/ncu_train.py tp_cfg4 4 > $O/ncu_train.log 2>&1; echo "ncu rc=$?"
python - <<P
import json, csv, collections
d=json.loads(open("$O/train_cfg4.json").read().strip().splitlines()[-1])
print("train", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["frac"])
agg=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(l for l in open("$O/train_glue_launches.csv") if l.startswith('"')):
    if r.get("Metric Name")=="gpu__time_duration.sum":
        k=r["Kernel Name"].split("(")[0][-40:]; v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
        v = v/1e3 if u in ("ns","nsecond") else v
        agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:42s} {n:5d} launches {t:10.1f} us  avg {t/n:8.1f}")
P
