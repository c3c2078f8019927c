#!/bin/bash
# GPU call: full gpu test suite, then ncu evidence (launch list of the bench command + one full capture of the fused kernel)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
# (1) launch list of the bench command (each launch once, gpu__time_duration only)
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
echo "launch list rc=$?"
# (2) full capture of the fused kernel on a short launch (65536 samples = 128 chunks)
timeout -s KILL 300 python tools/perf_probe.py 65536 4096 192 6 512 0 > gpurun_out/probe_plain.log 2>&1 &&
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/gram_full -f \
    python tools/perf_probe.py 65536 4096 192 6 512 0 > gpurun_out/probe_ncu.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -20
