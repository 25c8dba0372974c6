#!/bin/bash
# round 2, GPU call M (8 GPUs): config 4 on 8 GPUs with the final code (parity against one GPU, e2e with the zero-copy gather)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
echo "== c4 x8"; timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02m_bench_c4_8gpu.json 2> gpurun_out/r02m_bench_c4_8gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r02m_bench_c4_8gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02m_bench_c4_8gpu.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'value %.3e' % d['value'], 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('phases0', {k:round(v,3) for k,v in d['phase_ms_per_rank'][0].items()})
    print('parity', d.get('parity'))
    print('roof', round(d['roofline']['frac'],3), d['roofline'].get('whole_job'), d['roofline'].get('nvlink'))
except Exception as e:
    print('no line', e)
PY
