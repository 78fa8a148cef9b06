mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/gpu_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 2 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_v5.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],r['frac'],d['clocks'],d['selected_indices'])"
tail -5 gpurun_out/bench_v5.err
