#!/bin/bash
# round 2, GPU call E (8 GPUs): config 4 and config 5 (50M variable-length parts, k = 25) on 8 GPUs, parity against one GPU
set -u
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r02e_host.txt; nproc >> gpurun_out/r02e_host.txt; nvidia-smi -L >> gpurun_out/r02e_host.txt
avail=$(free -g | awk '/^Mem:/{print $7}')
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615"
echo "== c4 x8"; timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02e_bench_c4_8gpu.json 2> gpurun_out/r02e_bench_c4_8gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r02e_bench_c4_8gpu.err
if [ "$avail" -ge 220 ]; then
  echo "== c5 x8 (host memory available: ${avail} GB)"; timeout 1200 $TR bench.py --gpus 8 --config c5 --steps 3 --warmup 3 > gpurun_out/r02e_bench_c5_8gpu.json 2> gpurun_out/r02e_bench_c5_8gpu.err; echo "rc=$?"; tail -c 1500 gpurun_out/r02e_bench_c5_8gpu.err
else
  echo "== c5 x8 skipped: only ${avail} GB of host memory available"
fi
for f in c4_8gpu c5_8gpu; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02e_bench_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],3), 'value %.3e' % d['value'], 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('phases0', {k:round(v,3) for k,v in d['phase_ms_per_rank'][0].items()})
    print('parity', d.get('parity'))
    print('roof', round(d['roofline']['frac'],3), d['roofline']['whole_job'], d['roofline']['nvlink'])
    print('result', d['result'])
except Exception as e:
    print('no line', e)
PY
done
