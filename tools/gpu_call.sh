mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/gpu_tests.log
