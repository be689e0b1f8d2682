( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s6_tests.log 2>&1; tail -5 gpurun_out/s6_tests.log
python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.log; cut -c1-300 gpurun_out/s6_bench.json
python tools/bench_map.py --points 60000000 --out gpurun_out/s6_c4.json > /dev/null 2> gpurun_out/s6_c4.log; tail -2 gpurun_out/s6_c4.log; head -30 gpurun_out/s6_c4.json
python tools/bench_c3.py --scans 10000 --out gpurun_out/s6_c3.json > /dev/null 2> gpurun_out/s6_c3.log; tail -4 gpurun_out/s6_c3.log
