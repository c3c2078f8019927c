#!/bin/bash
# one GPU call: parity tests (bounded by timeout so a deadlocked persistent kernel cannot hang the box), smoke, bench
mkdir -p gpurun_out
free -g | head -2; nproc
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout -s KILL 1500 python bench.py ${BENCH_ARGS:---steps 1 --warmup 1} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
