mkdir -p gpurun_out
for w in cfg2 cfg1; do
  ncu --set full --clock-control none --import-source on -k regex:kws_fused_kernel -s 1 -c 1 -f -o gpurun_out/r02_fused_$w python tools/prof_fused.py $w --iters 1 > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log
done
ncu --set full --clock-control none --import-source on -k regex:kws_fused_kernel -s 3 -c 3 -f -o gpurun_out/r02_fused_cfg3 python tools/prof_fused.py cfg3 --iters 1 > gpurun_out/ncu_cfg3.log 2>&1; tail -1 gpurun_out/ncu_cfg3.log
ncu --set full --clock-control none --import-source on -k regex:kws_fused_kernel -s 1 -c 1 -f -o gpurun_out/r02_fused_cfg2_pool python tools/prof_fused.py cfg2 --iters 1 --pool-only > gpurun_out/ncu_cfg2_pool.log 2>&1; tail -1 gpurun_out/ncu_cfg2_pool.log
KWS_B200_LIB=$PWD/enhance-cb-whisper_b200/libkws_b200_dbg.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "channel_groups" 2>&1 | tail -2
