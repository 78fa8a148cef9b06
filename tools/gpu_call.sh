cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02_pytest_6.txt
timeout 900 python bench.py --configs > gpurun_out/r02_configs_c.jsonl 2> gpurun_out/r02_configs_c.err; echo configs rc=$?
tail -5 gpurun_out/r02_configs_c.err
cut -c1-420 gpurun_out/r02_configs_c.jsonl
python tools/c1_profile.py 2>&1 | tail -40 | cut -c1-180 | tee gpurun_out/r02_c1_profile_c.txt
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo bench rc=$?; tail -2 gpurun_out/r02_bench_quick.err
