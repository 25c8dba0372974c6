#!/bin/bash
# round 2, GPU call C (1 GPU): A/B of the wide filter variant (256 threads x 16 records), parity-checked
set -u
mkdir -p gpurun_out
for v in 0 1; do
  echo "== c2 filter_wide=$v"; timeout 600 python bench.py --steps 10 --warmup 3 --opt filter_wide=$v --api-parts 0 --no-cpu-baseline > gpurun_out/r02c_c2_fw$v.json 2> gpurun_out/r02c_c2_fw$v.err; echo "rc=$?"
done
echo "== c4 filter_wide=1"; timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --opt filter_wide=1 --api-parts 0 --no-cpu-baseline --no-verify > gpurun_out/r02c_c4_fw1.json 2> gpurun_out/r02c_c4_fw1.err; echo "rc=$?"
echo "== c1 filter_wide=1"; timeout 600 python bench.py --config c1 --steps 10 --warmup 3 --opt filter_wide=1 --api-parts 0 --no-cpu-baseline > gpurun_out/r02c_c1_fw1.json 2> gpurun_out/r02c_c1_fw1.err; echo "rc=$?"
for f in c2_fw0 c2_fw1 c4_fw1 c1_fw1; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02c_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'cand', d['result']['candidates'], d['result']['oversized'])
except Exception as e:
    print('no line', e)
PY
done
tail -n 3 gpurun_out/r02c_*.err
