cp enhance-cb-whisper_b200/libkws_b200.so /tmp/new.so
for wl in cfg1 cfg3; do
  for lib in new head new head; do
    if [ $lib = new ]; then cp /tmp/new.so enhance-cb-whisper_b200/libkws_b200.so; else cp enhance-cb-whisper_b200/libkws_b200_head.so enhance-cb-whisper_b200/libkws_b200.so; fi
    timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('$wl $lib', round(d['value']), d['roofline']['frac'], d['clocks']['sm_mhz'])"
  done
done 2>&1 | tee gpurun_out/ab_wg2.log
