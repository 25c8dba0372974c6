#!/bin/bash
# round 2, GPU call U (2 GPUs): the sharded parity cases and config 4 on 2 GPUs with the final kernels (level 2 without the digit
# array also runs in the owners' phase)
set -u
mkdir -p gpurun_out
echo "== pytest sharded (2 GPUs)"; timeout 1200 python -m pytest tests/test_gpu_sharded.py -q -m gpu > gpurun_out/r02u_pytest_sharded.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02u_pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
echo "== c4 x2"; timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02u_bench_c4_2gpu.json 2> gpurun_out/r02u_bench_c4_2gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r02u_bench_c4_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02u_bench_c4_2gpu.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('parity', d.get('parity',{}).get('vs_single_gpu'), 'single', d.get('single_gpu_same_config_ms'))
    print('phases', d.get('phase_ms_per_rank'))
except Exception as e:
    print('no line', e)
PY
