#!/bin/bash
# run on the GPU box: microbenchmarks that set the FP64 roofline denominator
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 250 > ../../gpurun_out/mb_clocks.csv &
SMI=$!
timeout 300 ./dmma_peak > ../../gpurun_out/mb_dmma.log 2>&1
timeout 300 python dgemm_peak.py > ../../gpurun_out/mb_dgemm.log 2>&1
kill $SMI
cat ../../gpurun_out/mb_dmma.log ../../gpurun_out/mb_dgemm.log
