"""profiles/roofline_traffic.json from `ncu --set full` captures of the shipped fused-kernel instances (no GPU needed).

    python tools/ncu_traffic.py cfg2=gpurun_out/r02_fused_cfg2.ncu-rep:148 cfg1=...:148 cfg3=...:148

Each entry: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) summed over the launches of ONE job of the capture
(multi-pass: all channel-group passes), the pairs that job scored, and the kernel instance(s) in the spelling bench.py
uses, so that bench.py only reports `roofline.traffic` when the capture is of the instance it actually ran.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def instance(name):  # "void kws_fused_kernel<1, 16, 0, 2, 1, 0>(...)" -> "kws_fused_kernel<1,16,0,2,1>" (RAGGED flag dropped)
    m = re.search(r"(kws_\w+)<([^>]*)>", name)
    if not m:
        return name.split("(")[0]
    args = [a.strip() for a in m.group(2).split(",")]
    return f"{m.group(1)}<{','.join(args[:5])}>"


def main():
    out = {}
    for spec in sys.argv[1:]:
        wl, rest = spec.split("=", 1)
        path, pairs = rest.rsplit(":", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        rd = wr = us = 0.0
        names = []
        for vals in rows[2:]:
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            f = lambda k: float(d[k][1].replace(",", "")) * UNIT.get(d[k][0], 1)
            rd += f("dram__bytes_read.sum")
            wr += f("dram__bytes_write.sum")
            us += float(d["gpu__time_duration.sum"][1].replace(",", ""))
            names.append(instance(d["Kernel Name"][1]))
        if len(names) == 1:
            kern = names[0]
        else:  # bench.py's spelling of the multi-pass chain
            first, last = names[0], names[-1]
            kern = f"{first} x{len(names) - 1} + {last[len('kws_fused_kernel'):]} (multi-pass: 12+12+8 layers)"
        out[wl] = {"kernel": kern, "pairs_per_launch": int(pairs), "dram_bytes_per_launch": rd + wr,
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "gpu_time_us_under_ncu": us,
                   "launches": names, "source": os.path.relpath(path, ROOT)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
