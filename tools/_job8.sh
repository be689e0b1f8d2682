( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s8_tests.log 2>&1; tail -5 gpurun_out/s8_tests.log
python tools/bench_map.py --points 60000000 --out gpurun_out/s8_c4.json > /dev/null 2> gpurun_out/s8_c4.log; tail -2 gpurun_out/s8_c4.log; python -c "
import json; d=json.load(open('gpurun_out/s8_c4.json')); print(d['points_per_s_device'], d['points_per_s_host_api'], d['frac_of_measured_hbm_whole_update'], d['counters']['n_evicted']); [print(r['kernel'], r['ms_total'], r['share'], r['frac_of_measured_hbm']) for r in d['per_kernel'][:8]]"
