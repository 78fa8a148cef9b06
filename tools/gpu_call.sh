cd /root/repo
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_chk.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['int8_tensor'], d['roofline']['fp64_pipe']['frac_of_issue_peak'])
PY
