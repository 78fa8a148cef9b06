mkdir -p gpurun_out
timeout 600 python tools/potentials_probe.py > gpurun_out/potentials_probe.log 2>&1; echo "rc=$?"
tail -12 gpurun_out/potentials_probe.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/gpu_tests.log
