set -u
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_glue_kernels_gpu.py tests/test_swin_gpu.py -m gpu -q -x > $O/pytest_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_taskprompter_gpu.py tests/test_invpt_gpu.py tests/test_big_goldens_gpu.py tests/test_module_forwards.py -m gpu -q > $O/pytest_models.log 2>&1; echo "models rc=$?"; tail -3 $O/pytest_models.log
for w in tp_cfg4 ip_cfg3 tps_swinB; do
  for gv in 1 0; do
    MTT_GEMM_GROUPED_VARIANT=$gv timeout 600 python bench.py --config $w --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > $O/bench_${w}_g$gv.json 2> $O/bench_${w}_g$gv.err
    python - <<P
import json
try:
    d=json.loads(open("$O/bench_${w}_g$gv.json").read().strip().splitlines()[-1]); print("$w", "g$gv", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("roofline",{}).get("frac"))
except Exception as e: print("$w g$gv FAIL", e)
P
  done
done
