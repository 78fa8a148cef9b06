set -x
mkdir -p gpurun_out
CMD="python bench.py --n 1000000 --steps 1 --warmup 1 --opt-itrs 2 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k k_project -c 3 -o gpurun_out/prof_project $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/ncu2.log
