timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3.json')); print('cfg3', round(d['value']), d['phases_ms'], d['roofline']['frac'])"
timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/bench_cfg2q.json 2> gpurun_out/bench_cfg2q.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2q.json')); print('cfg2', round(d['value']), d['roofline']['frac'])"
timeout 120 python tools/prof_kernels.py --only fused --pairs-k 148 --utts 8 --C 12 2>&1 | tail -1
