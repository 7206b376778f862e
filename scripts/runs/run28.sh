set -u
O=gpurun_out/r3i; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_taskprompter_gpu.py tests/test_invpt_gpu.py tests/test_swin_gpu.py tests/test_bench_cpu.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log | cut -c1-300
