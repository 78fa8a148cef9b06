cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_pytest_3.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; echo bench rc=$?
tail -3 gpurun_out/r02_bench_n1_a.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_a.json 2> gpurun_out/r02_bench_ref_a.err; echo ref rc=$?
tail -3 gpurun_out/r02_bench_ref_a.err
python tools/small_time.py 2>&1 | tail -30 | tee gpurun_out/r02_small_time_a.txt
python tools/c1_profile.py 2>&1 | tail -60 | cut -c1-180 | tee gpurun_out/r02_c1_profile_a.txt
