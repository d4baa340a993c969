timeout 300 python tools/body_probe.py 2>&1 | tail -12 | tee gpurun_out/body_probe2.log
timeout 300 python tools/body_probe.py --pairs 500 2>&1 | grep -E "^(eager|fused)" | tee -a gpurun_out/body_probe2.log
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 400 python bench.py --no-cpu --e2e-pairs 500 > gpurun_out/bench_e2e5.json 2> gpurun_out/bench_e2e5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e5.json')); print('e2e-pairs 500', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
