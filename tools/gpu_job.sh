KWS_FUSED_TIMERS=1 python enhance-cb-whisper_b200/build.py > /dev/null 2>&1
timeout 120 python tools/fused_trace.py 0 12 64 > gpurun_out/trace_cfg2.log 2>&1
timeout 120 python tools/fused_trace.py 0 4 384 > gpurun_out/trace_cfg1.log 2>&1
timeout 120 python tools/fused_trace.py 0 4 64 > gpurun_out/trace_c4.log 2>&1
head -30 gpurun_out/trace_cfg1.log
