"""Top stall-sample instructions of an `ncu --set full --import-source on` capture (SASS view; no GPU needed).

    python tools/ncu_hotspots.py gpurun_out/r02_mlp2.ncu-rep [N]

For warp-specialised kernels the sampled stalls show which role waits where: SYNCS.*TRYWAIT = mbarrier waits (which
barrier: see the surrounding code), LDG/LDS/STS = the data path of a role."""
import csv
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern = None
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        kern = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(dict(zip(hdr, rows[j])))
            j += 1
        tot = sum(int(r["# Samples"] or 0) for r in body)
        print(f"## {kern[:100]}: {tot} stall samples, {len(body)} SASS instructions")
        ranked = sorted(enumerate(body), key=lambda t: -int(t[1]["# Samples"] or 0))[:top]
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for idx, r in sorted(ranked):
            n = int(r["# Samples"] or 0)
            why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
            print(f"  [{idx:5d}] {100.0 * n / max(tot, 1):5.1f}%  {r['Source'].strip()[:70]:70s}  {', '.join(f'{w}:{c}' for c, w in why if c)}")
        i = j
    else:
        i += 1
