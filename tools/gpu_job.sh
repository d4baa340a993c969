timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "test_sim_stem_fused_matches_unfused_and_conv and (3-2-2 or 4-1-2)" 2>&1 | tail -3
cp enhance-cb-whisper_b200/libkws_b200.so /tmp/new.so
for lib in new head new head; do
  if [ $lib = new ]; then cp /tmp/new.so enhance-cb-whisper_b200/libkws_b200.so; else cp enhance-cb-whisper_b200/libkws_b200_head.so enhance-cb-whisper_b200/libkws_b200.so; fi
  timeout 200 python bench.py --workload cfg1 --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('cfg1 $lib', round(d['value']), d['roofline']['frac'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/ab_wg3.log
