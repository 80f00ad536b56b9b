#!/bin/bash
# Round evidence for the training step on one B200 (kept apart from capture_round.sh: gpurun_out/ carries <= 64 MiB per call).
# usage (under gpurun): bash tools/capture_train.sh r02w
tag=${1:-rXX}
out=gpurun_out
# training step: launch lists and one full capture each of the fused gradient kernel in lv and kl mode
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_lv_train.csv python tools/train_step.py --reps 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_kl_train.csv python tools/train_step.py --reps 2 --method kl > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lv_fused -s 1 -c 1 -o $out/${tag}_lv_fused python tools/train_step.py --reps 2 > $out/${tag}_ncu_lv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lv_fused -s 1 -c 1 -o $out/${tag}_kl_fused python tools/train_step.py --reps 2 --method kl > $out/${tag}_ncu_kl.log 2>&1
python tools/train_step.py --reps 8 > $out/${tag}_train_lv.txt 2>&1
python tools/train_step.py --reps 8 --method kl > $out/${tag}_train_kl.txt 2>&1
ls -la $out/${tag}_*
