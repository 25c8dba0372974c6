#!/bin/bash
# round 2, GPU call W (1 GPU): full-size parity of config 3 and the duplicate-heavy shape with the final code (level 2 without the
# digit array, smaller edge table), the reference arm
set -u
TAG=${1:-r02w}
mkdir -p gpurun_out
B="--api-parts 0 --no-cpu-baseline"
echo "== c3 --verify"; timeout 150 python bench.py --config c3 --verify --steps 5 --warmup 3 $B > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err; echo "rc=$?"
echo "== dup --verify"; timeout 150 python bench.py --config dup --verify --steps 3 --warmup 3 $B > gpurun_out/${TAG}_dup.json 2> gpurun_out/${TAG}_dup.err; echo "rc=$?"
echo "== reference arm"; timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "rc=$?"
for f in c3 dup; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), d.get('parity',{}).get('checked'))
except Exception as e:
    print('no line', e)
PY
done
tail -c 600 gpurun_out/${TAG}_ref.json
tail -n 3 gpurun_out/${TAG}_*.err
