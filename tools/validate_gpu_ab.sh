timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "sim_stem or channel_groups or cbw" 2>&1 | tail -2
for s in 0 1 0 1; do echo -n "cfg2-shape S12=$s: "; KWS_FUSED_S12=$s timeout 100 python tools/prof_kernels.py --only fused --pairs-k 148 --utts 8 --C 12 --iters 6 2>&1 | tail -1; done | tee gpurun_out/ab_s12.log
for s in 0 1; do KWS_FUSED_S12=$s timeout 150 python bench.py --workload cfg3 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ab.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('cfg3 S12=$s', round(d['value']), d['roofline']['frac'])"; done | tee -a gpurun_out/ab_s12.log
