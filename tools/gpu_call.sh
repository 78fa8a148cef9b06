cd /root/repo
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_b.json 2> gpurun_out/r02_bench_reference_b.err; echo ref rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo ncu rc=$?
python bench.py --configs > gpurun_out/r02_configs_l.jsonl 2> gpurun_out/r02_configs_l.err; echo configs rc=$?
