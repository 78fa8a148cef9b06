mkdir -p gpurun_out
timeout 300 python tools/q_probe.py > gpurun_out/q_probe.log 2>&1; echo "probe rc=$?"
tail -19 gpurun_out/q_probe.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/gpu_tests.log
