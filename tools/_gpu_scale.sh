N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r2_final_n$N.json 2> gpurun_out/bench_r2_final_n$N.err
tail -2 gpurun_out/bench_r2_final_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_final_n$N.json').read().strip().splitlines()[-1])
print('N=$N dev ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e value', round(d['e2e']['value']/1e6,1), 'launches', d['gpu_launches_per_step'])
print({k:(round(v['ms_per_step'],3), round(v['value']/1e6,1), v['launch_mode']) for k,v in d['other_configs'].items()})
PY
