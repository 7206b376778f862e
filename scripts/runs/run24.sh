set -u
O=gpurun_out/r3e; mkdir -p $O
for n in 1 2 4; do
  MTT_BILINEAR_PAIRS=$n timeout 300 python scripts/ncu_bilinear.py 2>&1 | head -1 | sed "s/^/pairs=$n: /" | tee -a $O/bilinear_pairs.txt
done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_glue_kernels_gpu.py -m gpu -q -k "bilinear" > $O/pytest_k.log 2>&1; echo "kernel tests rc=$?"; tail -2 $O/pytest_k.log | cut -c1-300
MTT_BILINEAR_PAIRS=1 timeout 600 python -m pytest tests/test_glue_kernels_gpu.py -m gpu -q -k "bilinear" > $O/pytest_k1.log 2>&1; echo "kernel tests (pairs=1) rc=$?"
MTT_BILINEAR_PAIRS=2 timeout 600 python -m pytest tests/test_glue_kernels_gpu.py -m gpu -q -k "bilinear" > $O/pytest_k2.log 2>&1; echo "kernel tests (pairs=2) rc=$?"
