cd /root/repo
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo bench rc=$?
python bench.py --configs > gpurun_out/r02_configs_k.jsonl 2> gpurun_out/r02_configs_k.err; echo configs rc=$?
python bench.py --n 1250000 --no-e2e --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r02_bench_shard_d.json 2> gpurun_out/r02_bench_shard_d.err; echo shard rc=$?
