#!/bin/bash
# ncu --set full capture of the fused kernel (one launch, 65536 samples at m=4096) and of one rollout step of the packed GEMM
mkdir -p gpurun_out
R=${ROUND:-r02}
export NK_NO_WARM=1
timeout -s KILL 300 python tools/perf_probe.py 65536 4096 192 6 0 0 > gpurun_out/${R}_probe_plain.log 2>&1 &&
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/${R}_gram_full -f \
    python tools/perf_probe.py 65536 4096 192 6 0 0 > gpurun_out/${R}_probe_ncu.log 2>&1
echo "gram capture rc=$?"; tail -2 gpurun_out/${R}_probe_plain.log
timeout -s KILL 300 python tools/rollout_probe.py 20000 4096 3 > gpurun_out/${R}_rollout_plain.log 2>&1 &&
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:pgemm_kernel -s 1 -c 1 -o gpurun_out/${R}_pgemm_full -f \
    python tools/rollout_probe.py 20000 4096 3 > gpurun_out/${R}_rollout_ncu.log 2>&1
echo "pgemm capture rc=$?"
