#!/bin/bash
# SM clock / power while one shipped fused-kernel shape runs back to back for a few seconds: is the shape power-capped?
#   tools/clock_probe.sh cfg2 cfg1 cfg3
for w in "$@"; do
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 100 > /tmp/clk_$w.csv &
  SM=$!
  timeout 120 python tools/prof_fused.py $w --pairs-k 592 --iters 400 2>&1 | tail -1
  kill $SM
  python - "$w" <<'PY'
import sys, statistics
w = sys.argv[1]
rows = [l.strip().split(", ") for l in open(f"/tmp/clk_{w}.csv") if l.strip()]
rows = rows[len(rows) // 3:]  # the kernel loop is running by then
mhz = [float(r[0].split()[0]) for r in rows]
pw = [float(r[2].split()[0]) for r in rows]
cap = sum(1 for r in rows if r[3].strip() == "Active")
print(f"  {w}: SM clock median {statistics.median(mhz):.0f} MHz of {rows[0][1]}, power median {statistics.median(pw):.0f} W (max {max(pw):.0f}), "
      f"sw_power_cap active in {cap}/{len(rows)} samples")
PY
done
