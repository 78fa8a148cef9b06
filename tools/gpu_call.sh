mkdir -p gpurun_out
timeout 300 python tools/dense_probe.py > gpurun_out/dense_probe.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/dense_probe.log
timeout 600 python -m pytest tests -m gpu -x -q -k "snnls or blackbox or hilbert" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log
