cd /root/repo
timeout 300 python tools/small_time.py 2>&1 | tail -9
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_final3.json 2> gpurun_out/bench_r01_final3.err
python - <<'PY'
import json,sys
d=json.loads(open('/root/repo/gpurun_out/bench_r01_final3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['host'], d['roofline']['launch_ms'], d['roofline']['share_of_step'], d['parity'])
PY
