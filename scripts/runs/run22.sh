set -u
O=gpurun_out/r3c; mkdir -p $O
timeout 300 python scripts/ncu_bilinear.py > $O/bilinear_timing.txt 2>&1; echo "timing rc=$?"; cat $O/bilinear_timing.txt | tail -5
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_glue_kernels_gpu.py -m gpu -q -k "bilinear or streamk" > $O/pytest_k.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/pytest_k.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bilinear_nhwc -c 2 -o $O/bilinear_nhwc python scripts/ncu_bilinear.py > $O/ncu_bilinear.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --config ip_cfg3 --steps 20 --warmup 5 --no-gpu-eager --no-cpu-baseline > $O/ip_cfg3.json 2>> $O/ab.err; echo "ip rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ip_cfg3.csv python scripts/ncu_forward.py ip_cfg3 parity 4 > $O/ncu_ip.log 2>&1; echo "ncu ip rc=$?"
python - <<P
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d.get("value"),1), round(d.get("ms_per_step"),3), (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e: print(f, "FAIL", e)
P
