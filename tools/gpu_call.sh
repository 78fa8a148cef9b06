cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo b200 rc=$?
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_last.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['parity'], {k:v['achieved_gbs'] for k,v in d['stage2']['kernels'].items()})
PY
