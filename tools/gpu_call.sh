cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r02_pytest_7.txt
timeout 900 python bench.py --configs > gpurun_out/r02_configs_d.jsonl 2> gpurun_out/r02_configs_d.err; echo configs rc=$?
tail -5 gpurun_out/r02_configs_d.err
cut -c1-420 gpurun_out/r02_configs_d.jsonl
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo bench rc=$?; tail -2 gpurun_out/r02_bench_quick.err
