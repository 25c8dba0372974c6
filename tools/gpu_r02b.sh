#!/bin/bash
# round 2, GPU call B (2 GPUs): sharded parity tests (new cases, shard-only staging, tensor gathers), C4 on 2 GPUs with parity,
# a scaled config 5 (variable lengths, k = 25) on 2 GPUs
set -u
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r02b_host.txt; nproc >> gpurun_out/r02b_host.txt; nvidia-smi -L >> gpurun_out/r02b_host.txt
echo "== pytest sharded"; timeout 1500 python -m pytest tests/test_gpu_sharded.py -q -m gpu -x > gpurun_out/r02b_pytest_sharded.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02b_pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
echo "== c4 x2"; timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02b_bench_c4_2gpu.json 2> gpurun_out/r02b_bench_c4_2gpu.err; echo "rc=$?"; tail -c 600 gpurun_out/r02b_bench_c4_2gpu.err
echo "== c5 scaled x2"; timeout 900 $TR bench.py --gpus 2 --config c5 --n-parts 4000000 --verify --steps 3 --warmup 3 > gpurun_out/r02b_bench_c5s_2gpu.json 2> gpurun_out/r02b_bench_c5s_2gpu.err; echo "rc=$?"; tail -c 1200 gpurun_out/r02b_bench_c5s_2gpu.err
for f in c4_2gpu c5s_2gpu; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('phases', [{k:round(v,3) for k,v in p.items()} for p in d['phase_ms_per_rank']])
    print('parity', d.get('parity'))
    print('roof', round(d['roofline']['frac'],3), d['roofline']['whole_job'], d['roofline']['nvlink'])
except Exception as e:
    print('no line', e)
PY
done
