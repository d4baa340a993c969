for v in "12,2,2,1,0" "12,2,2,1,1" "12,2,2,1,0" "12,2,2,1,1"; do
  KWS_FUSED_MULTI=$v timeout 200 python bench.py --workload cfg3 --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('$v', round(d['value']), d['phases_ms']['pairs_ms'], d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/ab_order.log
