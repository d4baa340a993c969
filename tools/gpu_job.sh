for c in 8 6; do for ns in 0 4 0 4; do
  echo -n "C=$c NS=$ns: "; KWS_FUSED_NS=$ns timeout 120 python tools/prof_kernels.py --only fused --pairs-k 148 --utts 8 --C $c --iters 5 2>&1 | tail -1
done; done | tee gpurun_out/ab_ns.log
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "test_sim_stem_fused_matches_unfused_and_conv" 2>&1 | tail -2
KWS_FUSED_NS=4 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "test_sim_stem_fused_matches_unfused_and_conv or channel_groups" 2>&1 | tail -2
