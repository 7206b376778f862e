set -u
O=gpurun_out/r3d; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_glue_kernels_gpu.py -m gpu -q -k "bilinear" > $O/pytest_k.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/pytest_k.log | cut -c1-300
timeout 300 python scripts/ncu_bilinear.py > $O/bilinear_timing.txt 2>&1; echo "timing rc=$?"; cat $O/bilinear_timing.txt | tail -5
timeout 900 python -m pytest tests/test_big_goldens_gpu.py tests/test_invpt_gpu.py -m gpu -q -k "cfg4_b4 or ip_" > $O/pytest_big.log 2>&1; echo "goldens rc=$?"; tail -2 $O/pytest_big.log | cut -c1-300
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ip_cfg3.csv python scripts/ncu_forward.py ip_cfg3 parity 4 > $O/ncu_ip.log 2>&1; echo "ncu ip rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg4.csv python scripts/ncu_forward.py tp_cfg4 parity 4 > $O/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 300 python bench.py --config ip_cfg3 --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline > $O/ip_cfg3.json 2>> $O/ab.err; echo "ip rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline --no-train-leg > $O/cfg4.json 2>> $O/ab.err; echo "cfg4 rc=$?"
grep -E "bilinear" $O/launches_ip_cfg3.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' | sort | uniq -c | sort -rn | head -5
python - <<P
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d.get("value"),1), round(d.get("ms_per_step"),3), (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e: print(f, "FAIL", e)
P
