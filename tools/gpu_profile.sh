#!/bin/bash
# Round profile recipe (run on the GPU box through gpurun): bench line of the driver's command, reference arm, ncu launch list and
# one ncu --set full capture of the timed steps of the resident pass (bench.py --ncu-range brackets them with
# cudaProfilerStart/Stop).  Outputs under gpurun_out/; summarise here with tools/ncu_summary.py.
#   usage: tools/gpu_profile.sh <tag> [workload]      (SKIP_BENCH=1: profiles only, SKIP_FULL=1: no --set full capture)
set -u
TAG=${1:-r02}; WL=${2:-c2}
mkdir -p gpurun_out
if [ "${SKIP_BENCH:-0}" != "1" ]; then
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; tail -2 gpurun_out/${TAG}_bench.log; cut -c1-400 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.log; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
fi
# launch list of the timed steps (cold-cache, serialised: compare SHARES with the live per-kernel numbers of the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --workload $WL --steps 8 --warmup 3 --no-cpu --no-c1 --no-loops --ncu-range > gpurun_out/${TAG}_ncu_launch.log 2>&1
if [ "${SKIP_FULL:-0}" != "1" ]; then
# one scan of the resident pass with the full metric set (a report with source counters is ~2 MB per launch; gpurun brings back <= 64 MiB)
ncu --set full --clock-control none --import-source on --profile-from-start off -c ${FULL_COUNT:-20} -o gpurun_out/${TAG}_prof -f \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-c1 --no-loops --ncu-range > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full.csv 2>/dev/null
fi
ls -la gpurun_out | tail -8
