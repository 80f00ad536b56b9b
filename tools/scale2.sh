python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_1gpu.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r02i_bench_2gpu.json 2>gpurun_out/r02i_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --scaling strong --global-batch 65536 > gpurun_out/r02i_bench_2gpu_strong64k.json 2>gpurun_out/r02i_bench_2gpu_strong.err
for f in 1gpu 2gpu 2gpu_strong64k; do python -c "
import json,sys;d=json.load(open('gpurun_out/r02i_bench_$f.json'));print('$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'])"; done
