cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_25.txt 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02_pytest_25.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err; echo bench rc=$?
python bench.py --configs > gpurun_out/r02_configs_n.jsonl 2> gpurun_out/r02_configs_n.err; echo configs rc=$?
python tools/small_time.py > gpurun_out/r02_small_time_b.txt 2>&1
