set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
timeout 120 python tools/prof_kernels.py --only rows,mlp,temporal,maxpool --pairs-k 250 --utts 1 > gpurun_out/prof_proj.log 2>&1; cat gpurun_out/prof_proj.log
timeout 120 python tools/prof_kernels.py --only rows,mlp,temporal --pairs-k 64 --utts 1 --C 32 --D 1280 >> gpurun_out/prof_proj.log 2>&1; tail -4 gpurun_out/prof_proj.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'rows_kernel|kws_gemm_kernel|temporal_kernel|maxpool' -s 8 -c 10 -o gpurun_out/prof_proj_r01 -f python tools/prof_kernels.py --only rows,mlp,temporal,maxpool --pairs-k 250 --utts 1 --iters 1 > gpurun_out/ncu_proj.log 2>&1; tail -3 gpurun_out/ncu_proj.log
