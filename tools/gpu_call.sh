mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 2 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_v6.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],r['share_of_step'],d['host'],d['selected_indices'])"
tail -5 gpurun_out/bench_v6.err
