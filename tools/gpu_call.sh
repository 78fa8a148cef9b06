cd /root/repo
timeout 300 python tools/small_time.py 2>&1 | tail -9
for s in newton hybrid; do
timeout 900 python bench.py --steps 3 --warmup 3 --sampler $s --no-cpu-baseline > gpurun_out/bench_sampler_$s.json 2> gpurun_out/bench_sampler_$s.err
python - $s <<'PY'
import json,sys
d=json.loads(open('/root/repo/gpurun_out/bench_sampler_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['host'], d['roofline']['launch_ms'], d['roofline']['share_of_step'])
PY
done
