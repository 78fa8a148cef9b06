mkdir -p gpurun_out
timeout 300 python tools/q_probe.py > gpurun_out/q_probe.log 2>&1; echo "probe rc=$?"
tail -6 gpurun_out/q_probe.log
