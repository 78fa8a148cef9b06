cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo ref rc=$?
timeout 900 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; echo b200 rc=$?
python - <<'PY'
import json
r=json.loads(open('/root/repo/gpurun_out/bench_ref_final.json').read().strip().splitlines()[-1])
d=json.loads(open('/root/repo/gpurun_out/bench_default_final.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['cores'], r['ms_per_step'])
print('b200', {k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline']['value'], d['parity'], d['clocks'])
PY
