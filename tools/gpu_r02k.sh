#!/bin/bash
# round 2, GPU call K (2 GPUs): whole GPU suite incl. the sharded cases after the kernel changes of this session; c4 on 2 GPUs with parity
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all, 2 GPUs)"; timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02k_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02k_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
echo "== c4 x2"; timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02k_bench_c4_2gpu.json 2> gpurun_out/r02k_bench_c4_2gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r02k_bench_c4_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02k_bench_c4_2gpu.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('parity', d.get('parity',{}).get('vs_single_gpu'), 'single', d.get('single_gpu_same_config_ms'))
except Exception as e:
    print('no line', e)
PY
