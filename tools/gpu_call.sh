cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/q_time.py 2>&1 | tail -1 | tee gpurun_out/r02_q_tab_e.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_21.txt 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02_pytest_21.txt
python bench.py > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; echo bench rc=$?
