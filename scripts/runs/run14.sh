set -u
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/ncu_train.py tp_cfg4 4 > $O/ncu_train.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_train.log
python scripts/summarize_ncu.py --launches $O/train_launches.csv $O/train_launch_shares.md; head -45 $O/train_launch_shares.md
