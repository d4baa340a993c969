set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 120 python tools/prof_kernels.py --only mlp,temporal --pairs-k 250 --utts 1 > gpurun_out/prof_proj2.log 2>&1
timeout 120 python tools/prof_kernels.py --only mlp,temporal --pairs-k 64 --utts 1 --C 32 --D 1280 >> gpurun_out/prof_proj2.log 2>&1; cat gpurun_out/prof_proj2.log
timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -2 gpurun_out/bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3.json')); print(d['value'], d['phases_ms'], d['roofline']['frac'], d['hbm'], d['projection_tflops'])"
