# Round 2: ncu launch list of two training steps (forward + backward, batch 64 x 1024).  bash profiles/session_j.sh
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_r02_training.csv \
    python benchmarks/profile_training_step.py 64 --plain 2 > $O/launches_r02_training.log 2>&1
tail -2 $O/launches_r02_training.log; wc -l $O/launches_r02_training.csv
