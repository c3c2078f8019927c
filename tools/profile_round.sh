#!/bin/bash
# GPU call: ncu evidence for the round (each capture only after the same command exited 0 without ncu)
mkdir -p gpurun_out
R=${ROUND:-r02}
export NK_NO_WARM=1     # no engine warm-up launches in front of the captured kernels
# (1) launch list of the bench command
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-config5 > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err &&
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-config5 > gpurun_out/${R}_bench_ncu.json 2> gpurun_out/${R}_bench_ncu.err
echo "launch list rc=$?"
# (2) full capture of the fused kernel on a short launch (65536 samples = 128 chunks)
timeout -s KILL 300 python tools/perf_probe.py 65536 4096 192 6 512 0 > gpurun_out/${R}_probe_plain.log 2>&1 &&
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/${R}_gram_full -f \
    python tools/perf_probe.py 65536 4096 192 6 512 0 > gpurun_out/${R}_probe_ncu.log 2>&1
echo "gram capture rc=$?"
# (3) full capture of the packed persistent GEMM (one rollout step, 20000 trajectories, m=4096)
timeout -s KILL 300 python tools/rollout_probe.py 20000 4096 3 > gpurun_out/${R}_rollout_plain.log 2>&1 &&
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:pgemm_kernel -s 1 -c 1 -o gpurun_out/${R}_pgemm_full -f \
    python tools/rollout_probe.py 20000 4096 3 > gpurun_out/${R}_rollout_ncu.log 2>&1
echo "pgemm capture rc=$?"
# (4) full capture of the TMA-fed dense GEMM (Newton-Schulz product of the symmetric square root, m=4096)
timeout -s KILL 300 python tools/dense_probe.py 4096 > gpurun_out/${R}_dense_plain.log 2>&1 &&
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 4 -c 1 -o gpurun_out/${R}_tgemm_full -f \
    python tools/dense_probe.py 4096 > gpurun_out/${R}_dense_ncu.log 2>&1
echo "tgemm capture rc=$?"
cat gpurun_out/${R}_rollout_plain.log | tail -2
ls -la gpurun_out | tail -12
