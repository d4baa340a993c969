"""Summarise ncu output for profiles/ (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [--json profiles/roofline_traffic.json
                                --workload cfg2 --pairs 1184] > profiles/rNN_fused_full.md
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_uniform", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__sass_inst_executed_op_shared_st.sum",
    "smsp__cycles_active.avg", "sm__ctas_launched.sum",
    # the counters that DO track tcgen05.mma occupancy on sm_100 (the *_realtime variants read "no data" / far too low)
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    # L2 -> SM fabric: what the fused kernel's operand pipeline is bound by
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex.sum",
    "lts__t_sectors_srcunit_tex.sum",
]


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(d["Metric Unit"], v)
        a = agg.setdefault(d["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")


def full(path, extra):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for li, vals in enumerate(rows[2:]):
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print(f"### launch {li}: `{d.get('Kernel Name', ('', '?'))[1][:80]}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    print(f"| {h} | {d[h][1]} | {d[h][0]} |")
                    break
        print()
        res.append(d)
    if "--json" in extra:
        jp = extra[extra.index("--json") + 1]
        wl = extra[extra.index("--workload") + 1]
        pairs = int(extra[extra.index("--pairs") + 1])
        d = res[0]

        def by(name):
            u, v = d[name]
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

        tr = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
        json.dump({"workload": wl, "kernel": d["Kernel Name"][1].split("(")[0], "pairs_per_launch": pairs,
                   "dram_bytes_per_launch": tr, "dram_bytes_read": by("dram__bytes_read.sum"),
                   "dram_bytes_write": by("dram__bytes_write.sum"), "source": path}, open(jp, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3:])
