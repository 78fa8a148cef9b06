cd /root/repo
timeout 600 python -m pytest tests/test_gpu_sampler.py -x -q 2>&1 | tail -5
for s in hybrid newton; do
timeout 900 python bench.py --steps 3 --warmup 3 --sampler $s --no-cpu-baseline > gpurun_out/bench_sampler_$s.json 2> gpurun_out/bench_sampler_$s.err
python - $s <<'PY'
import json,sys
d=json.loads(open('/root/repo/gpurun_out/bench_sampler_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['host'])
PY
done
