"""bench.py contract pieces that can be checked without a GPU: workload table vs BASELINE.md section 5,
argument defaults, the JSON keys of the reference arm (run on a tiny budget)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_work_per_pair_matches_baseline_md():
    sys.path.insert(0, ROOT)
    import bench

    # GFLOP per pair (similarity + stem)
    for name, sim, stem in (("cfg1", 0.691, 1.411), ("cfg2", 0.346, 4.234), ("cfg3", 0.230, 2.860)):
        s, t = bench.flops_per_pair(bench.WORKLOADS[name])
        assert abs(s / 1e9 - sim) < 2e-3 and abs(t / 1e9 - stem) < 2e-3, (name, s, t)
    wl = bench.WORKLOADS["cfg2"]
    assert (wl["K"], wl["U"], wl["C"], wl["D"], wl["P"], wl["Tk"], wl["Tu"]) == (1000, 256, 12, 768, 64, 150, 1500)
    # projection: SURVEY 8d, 4.09 TFLOP in total at cfg2
    assert abs(bench.flops_projection(wl, wl["K"], wl["U"]) / 1e12 - 4.09) < 0.05


def test_issued_flops_account_for_the_kernels_packing():
    """roofline.issued: every tcgen05.mma of the fused kernel at 2 M N K.  cfg2: 13 column tiles x 38 steps x 21 stem
    MMAs of 128 x 128 x 16 and 10 chunks x 12 layers x 4 similarity MMAs of 128 x 16 x 16 per tile."""
    sys.path.insert(0, ROOT)
    import bench

    sim, stem = bench.issued_flops_per_pair(bench.WORKLOADS["cfg2"])
    assert stem == 13 * 38 * 21 * 2.0 * 128 * 128 * 16
    assert sim == 13 * 10 * 12 * 4 * 2.0 * 128 * 16 * 16
    for name in ("cfg1", "cfg2", "cfg3"):
        wl = bench.WORKLOADS[name]
        eff = sum(bench.flops_per_pair(wl)) / sum(bench.issued_flops_per_pair(wl))
        assert 0.70 < eff < 0.80, (name, eff)  # 60 of 64 pixel slots, 84 of 96 tap slots, partly empty last tile
    # cfg1 (4 layers): ONE stem MMA per kernel row; cfg3: passes of 12 + 12 + 8 layers (3 + 3 + 2 MMAs per kernel row)
    assert bench.issued_flops_per_pair(bench.WORKLOADS["cfg1"])[1] == 13 * 38 * 7 * 2.0 * 128 * 128 * 16
    assert bench.issued_flops_per_pair(bench.WORKLOADS["cfg3"])[1] == 7 * 19 * 7 * 8 * 2.0 * 128 * 128 * 16


def test_b200_arm_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "1", "--ref-budget", "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "kwd_utt_pairs_per_s" and line["unit"] == "pairs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    # like-for-like scopes (VERDICT r1 / ADVICE r1): `value` is the in-scope figure (the scope of the B200 arm's `value`),
    # `e2e.value` the whole forward through logits (the scope of the B200 arm's `e2e.value`), which is slower
    assert line["cpu_baseline"]["value"] == line["value"] and line["config"]["same_scope_as_b200_value"] is True
    assert line["e2e"]["unit"] == "pairs/s" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert 0 < line["e2e"]["value"] < line["value"]
