#!/bin/bash
# e2e throughput vs the body's chunk size (pairs per cuDNN batch): tools/e2e_sweep.sh 250 500 1000 2000
for ep in "$@"; do
  timeout 300 python bench.py --only e2e --no-parity --no-cpu --utts 64 --e2e-pairs $ep --e2e-steps 2 2>/dev/null > /tmp/e2e_$ep.json
  python - "$ep" <<'PY'
import json, sys
ep = sys.argv[1]
d = json.load(open(f"/tmp/e2e_{ep}.json"))
print(f"e2e-pairs {ep}: {d['e2e']['value']:.0f} pairs/s, {d['e2e']['ms_per_step']:.1f} ms per step", flush=True)
PY
done
