mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_final_n1.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],r['share_of_step'],d['e2e']['value'],d['cpu_baseline']['value'],d['parity'],d['host'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/bench_final_ref.json
CMD="python bench.py --rows 1000000 --steps 1 --warmup 1 --opt-itrs 2 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_project_q -c 2 -f -o gpurun_out/prof_bench_final $CMD > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc=$?"
