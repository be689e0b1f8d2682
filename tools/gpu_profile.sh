#!/bin/bash
# Round profile recipe (run on the GPU box through gpurun): bench line, ncu launch list of the same command, one ncu --set full
# capture of two scans of the resident pass.  Outputs under gpurun_out/; summarise here with tools/ncu_summary.py.
#   usage: tools/gpu_profile.sh <tag> [workload]
set -u
TAG=${1:-r01}; WL=${2:-c2}
mkdir -p gpurun_out
if [ "${SKIP_BENCH:-0}" != "1" ]; then
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; tail -2 gpurun_out/${TAG}_bench.log; cut -c1-600 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.log; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-300
fi
# launch list: 8 steps + 3 warm-up; ~20 kernels per scan, the resident pass (graph replay) follows the end-to-end pass (13 scans)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 290 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-cpu --no-c1 --no-loops > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 330 -c 44 -o gpurun_out/${TAG}_prof -f \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-cpu --no-c1 --no-loops > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
