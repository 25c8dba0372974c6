#!/bin/bash
# round 2, GPU call G (2 GPUs): walk grouping + e2e gather on 2 GPUs; full GPU suite; single-GPU bench A/B of the grouping variants
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r02g_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02g_pytest.log
for v in 1 0; do
  echo "== c2 group_walk=$v"; timeout 600 python bench.py --steps 10 --warmup 3 --opt group_walk=$v --api-parts 0 --no-cpu-baseline > gpurun_out/r02g_c2_gw$v.json 2> gpurun_out/r02g_c2_gw$v.err; echo "rc=$?"
done
echo "== c1"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --api-parts 0 --no-cpu-baseline > gpurun_out/r02g_c1.json 2> gpurun_out/r02g_c1.err; echo "rc=$?"
echo "== c3"; timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --api-parts 0 --no-cpu-baseline --verify > gpurun_out/r02g_c3.json 2> gpurun_out/r02g_c3.err; echo "rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29619"
echo "== c4 x2"; timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02g_bench_c4_2gpu.json 2> gpurun_out/r02g_bench_c4_2gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r02g_bench_c4_2gpu.err
for f in c2_gw1 c2_gw0 c1 c3; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02g_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'launches', d['gpu_launches'])
except Exception as e:
    print('no line', e)
PY
done
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02g_bench_c4_2gpu.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('parity', d.get('parity',{}).get('vs_single_gpu'))
except Exception as e:
    print('no line', e)
PY
