#!/bin/bash
# Round evidence on one B200: bench lines of every workload, launch list and one `ncu --set full` capture of the rollout kernel.
# usage (under gpurun): bash tools/capture_round.sh r02m
tag=${1:-rXX}
out=gpurun_out
for w in gmm50 cfg2 cfg3 cfg4 gmm50dense; do
  python bench.py --steps 30 --warmup 3 --workload $w > $out/${tag}_bench_$w.json 2>$out/${tag}_bench_$w.err
done
python bench.py --workload cfg5 > $out/${tag}_bench_cfg5.json 2>$out/${tag}_bench_cfg5.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 6 -c 1 -o $out/${tag}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/${tag}_clocks.csv
ls -la $out/${tag}_*
