# Round 2, last session: full verification (session_c) + kernel profile and ncu launch list of the training step.
set -u
O=gpurun_out; mkdir -p $O
bash profiles/session_c.sh
timeout 300 python benchmarks/profile_training_step.py > $O/r02_training_step_profile.txt 2>&1; head -12 $O/r02_training_step_profile.txt | cut -c1-150
bash profiles/session_j.sh
