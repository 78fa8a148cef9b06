mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r01_q.json 2> gpurun_out/bench_r01_q.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_r01_q.json; tail -3 gpurun_out/bench_r01_q.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref2.json 2> gpurun_out/bench_r01_ref2.err; echo "ref rc=$?"
cat gpurun_out/bench_r01_ref2.json | cut -c1-600
CMD="python bench.py --n 1000000 --steps 1 --warmup 1 --opt-itrs 2 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_q.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_project_q -c 2 -f -o gpurun_out/prof_bench_q $CMD > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc=$?"
