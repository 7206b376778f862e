set -u
O=gpurun_out/r2l; mkdir -p $O
MTT_TRAIN_TEST_VERBOSE=1 timeout 1500 python -m pytest tests/test_train_gpu.py -m gpu -q -s -k "training_step or torch_facing" > $O/pytest_train.log 2>&1; echo "train rc=$?"; grep -E "worst|passed|failed|Error|assert " $O/pytest_train.log | cut -c1-1600 | tail -30
