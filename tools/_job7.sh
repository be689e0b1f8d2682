( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s7_tests.log 2>&1; tail -5 gpurun_out/s7_tests.log
python tools/bench_c3.py --scans 3600 --out gpurun_out/s7_c3.json > /dev/null 2> gpurun_out/s7_c3.log; tail -12 gpurun_out/s7_c3.log | cut -c1-300
python tools/bench_map.py --points 60000000 --out gpurun_out/s7_c4.json > /dev/null 2> gpurun_out/s7_c4.log; tail -2 gpurun_out/s7_c4.log; grep -E "points_per_s|k_lru_evict|frac_of_measured_hbm_whole" -A1 gpurun_out/s7_c4.json | head -20
python bench.py --no-c1 --no-cpu --no-loops --steps 40 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.log; cut -c1-200 gpurun_out/s7_bench.json; python -c "
import json; d=json.load(open('gpurun_out/s7_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['per_kernel_us_per_step'])"
