#!/bin/bash
# round 2, GPU call D (2 GPUs): the whole GPU suite (single-GPU + sharded tests), config 4 on 2 GPUs (sparse flags, page-locked gathers)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r02d_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02d_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
echo "== c4 x2"; timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_bench_c4_2gpu.json 2> gpurun_out/r02d_bench_c4_2gpu.err; echo "rc=$?"; tail -c 400 gpurun_out/r02d_bench_c4_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench_c4_2gpu.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('phases', [{k:round(v,3) for k,v in p.items()} for p in d['phase_ms_per_rank']])
    print('parity', d.get('parity',{}).get('vs_single_gpu'), d.get('single_gpu_same_config_ms'))
except Exception as e:
    print('no line', e)
PY
