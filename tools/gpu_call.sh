set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/gpu_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_v2.json; tail -5 gpurun_out/bench_v2.err
CMD="python bench.py --n 1000000 --steps 1 --warmup 1 --opt-itrs 2 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k k_project -c 2 -o gpurun_out/prof_project_v2 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
