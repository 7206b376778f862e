set -u
O=gpurun_out/r2k; mkdir -p $O
timeout 1500 python -m pytest tests/test_train_gpu.py -m gpu -q -x --deselect "tests/test_train_gpu.py::test_training_step_matches_reference[tp_cfg4_d4]" > $O/pytest_train.log 2>&1; echo "train rc=$?"; tail -40 $O/pytest_train.log
MTT_TRAIN_TEST_VERBOSE=1 timeout 900 python -m pytest "tests/test_train_gpu.py::test_training_step_matches_reference[tp_cfg4_d4]" -m gpu -q -x -s > $O/pytest_train_cfg4d4.log 2>&1; echo "cfg4d4 rc=$?"; tail -30 $O/pytest_train_cfg4d4.log | cut -c1-1500
