set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r01_n1.json; tail -5 gpurun_out/bench_r01_n1.err
python bench.py --impl reference > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; echo "ref rc=$?"
python bench.py --n 2000000 --steps 1 --warmup 1 --opt-itrs 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --n 2000000 --steps 1 --warmup 1 --opt-itrs 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
