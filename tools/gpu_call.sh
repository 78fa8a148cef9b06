mkdir -p gpurun_out
timeout 600 python tools/debug_c1.py 30 > gpurun_out/debug_c1.log 2>&1; echo "rc=$?"
tail -12 gpurun_out/debug_c1.log
