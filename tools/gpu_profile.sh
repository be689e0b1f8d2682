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
# launch list of the resident pass of a shorter run of the same command (24 steps + 6 warm-up: ~20 kernels per scan, the
# end-to-end pass of 32 scans comes first, the timed region of the resident pass starts 6 scans later), plus the line of that
# same command without ncu, so that the shares can be compared
python bench.py --workload $WL --steps 24 --warmup 6 --no-cpu --no-c1 --no-loops > gpurun_out/${TAG}_bench_24.json 2>> gpurun_out/${TAG}_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 775 -c 480 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --workload $WL --steps 24 --warmup 6 --no-cpu --no-c1 --no-loops > gpurun_out/${TAG}_ncu_launch.log 2>&1
# one scan of the resident pass with the full metric set (no source import: the report has to stay small enough to travel)
ncu --set full --clock-control none -k regex:k_ -s 300 -c 22 -o gpurun_out/${TAG}_prof -f \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-cpu --no-c1 --no-loops > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
