#!/bin/bash
# round 2, GPU call F (4 GPUs): the sharded parity tests on 4 GPUs (log kept under profiles/), config 4 on 4 GPUs with e2e phases
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_host.txt
echo "== pytest sharded (4 GPUs)"; timeout 1800 python -m pytest tests/test_gpu_sharded.py -v -m gpu -x > gpurun_out/r02f_pytest_sharded_4gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02f_pytest_sharded_4gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29617"
echo "== c4 x4"; timeout 900 $TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02f_bench_c4_4gpu.json 2> gpurun_out/r02f_bench_c4_4gpu.err; echo "rc=$?"; tail -c 400 gpurun_out/r02f_bench_c4_4gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02f_bench_c4_4gpu.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('phases0', {k:round(v,3) for k,v in d['phase_ms_per_rank'][0].items()})
    print('parity', d.get('parity',{}).get('vs_single_gpu'), d.get('single_gpu_same_config_ms'))
except Exception as e:
    print('no line', e)
PY
