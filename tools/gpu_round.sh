#!/bin/bash
# one GPU call: parity tests (bounded by timeout so a deadlocked persistent kernel cannot hang the box) + perf probe
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/perf_probe.py ${PROBE_N:-400000} 4096 192 6 512 2 2>&1 | tee gpurun_out/perf_probe.log
