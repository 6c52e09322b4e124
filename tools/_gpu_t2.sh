mkdir -p gpurun_out
(timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) 
for c in 1 0; do
  PROMP_B200_CHAIN=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain$c.json 2> gpurun_out/bench_chain$c.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_chain$c.json').read().strip().splitlines()[-1])
print('chain=$c', 'dev ms', round(d['ms_per_step'],4), 'eager', round(d['eager_ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'e2e eager', round(d['e2e']['eager']['ms_per_step'],4), 'launches', d['gpu_launches_per_step'])
print({k:(round(v['ms_per_step'],3)) for k,v in d['other_configs'].items()})
print({k:(v['launches_per_iter'], round(v['avg_ms']*1e3,1)) for k,v in d['kernels'].items()})
PY
done
