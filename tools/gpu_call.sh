mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/gpu_tests.log
timeout 1200 python bench.py > gpurun_out/bench_r01_q7.json 2> gpurun_out/bench_r01_q7.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_r01_q7.json'));r=d['roofline'];print(d['value'],d['ms_per_step'],r['launch_ms'],r['score_pass_ms'],r['share_of_step'],d['e2e']['value'],d['cpu_baseline']['value'],d['parity']);print(json.dumps(d['stage2'],indent=0)[:900]);print(d['stage3'])"
tail -5 gpurun_out/bench_r01_q7.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
