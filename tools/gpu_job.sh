timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 200 python tools/cfg4_bench.py --K 1000 --S 16 2>&1 | tail -2
timeout 200 python tools/cfg4_bench.py --K 1000 --S 16 --unfused 2>&1 | tail -2
