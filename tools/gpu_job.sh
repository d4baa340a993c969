timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_fused.py enhance-cb-whisper_b200/libkws_b200_head.so enhance-cb-whisper_b200/libkws_b200.so enhance-cb-whisper_b200/libkws_b200_head.so enhance-cb-whisper_b200/libkws_b200.so 2>&1 | tee gpurun_out/ab_mailbox.log | tail -12
