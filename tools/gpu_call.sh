cd /root/repo
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1_h.json 2> gpurun_out/r02_bench_n1_h.err; echo bench rc=$?
