#!/bin/bash
# weak and strong scaling lines on N GPUs of one box: bash tools/scale_n.sh <N> <tag>
N=${1:-8}; tag=${2:-rXX}; out=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 30 --warmup 3 "${@:3}" > $out/${tag}_bench_${N}gpu_$2.json 2>$out/${tag}_bench_${N}gpu_$2.err; }
run 29521 weak
run 29522 strong64k --scaling strong --global-batch 65536
run 29523 strong256k --scaling strong --global-batch 262144
run 29524 cfg5 --workload cfg5
for f in weak strong64k strong256k cfg5; do python -c "
import json;d=json.load(open('$out/${tag}_bench_${N}gpu_$f.json'));print('$N gpus $f',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline'].get('kernel_ms'))"; done
