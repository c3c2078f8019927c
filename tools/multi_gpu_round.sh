#!/bin/bash
# N-GPU check of the sharded fit: bench.py under torchrun at reduced n (quick) and then the named n
mkdir -p gpurun_out
N=${NGPU:-2}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 1 --warmup 1 --samples 1000000 --no-cpu > gpurun_out/bench_${N}gpu_small.json 2> gpurun_out/bench_${N}gpu_small.err
echo "rc=$?"; tail -3 gpurun_out/bench_${N}gpu_small.err; cat gpurun_out/bench_${N}gpu_small.json | cut -c1-1500
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err; cat gpurun_out/bench_${N}gpu.json | cut -c1-2500
