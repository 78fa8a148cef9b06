cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_16.txt 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_pytest_16.txt
python bench.py --n 1250000 --no-e2e --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r02_bench_shard_c.json 2> gpurun_out/r02_bench_shard_c.err; echo shard rc=$?
python bench.py --n 1250000 --no-e2e --no-cpu-baseline --steps 1 --warmup 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r02_shard_launches_c.csv python bench.py --n 1250000 --no-e2e --no-cpu-baseline --steps 1 --warmup 2 > gpurun_out/ncu_shard_c.log 2>&1; echo ncu rc=$?
python bench.py --n 1250000 --no-e2e --no-cpu-baseline --steps 1 --warmup 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_project_q -s 30 -c 2 -f -o gpurun_out/r02_q_full python bench.py --n 1250000 --no-e2e --no-cpu-baseline --steps 1 --warmup 2 > gpurun_out/ncu_q_full.log 2>&1; echo ncu full rc=$?
