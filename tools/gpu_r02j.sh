#!/bin/bash
# round 2, GPU call J (1 GPU): whole GPU suite (general path, packed entry, folded flags), c2 with / without the overlapped tail, c1
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02j_pytest.log
B="--api-parts 0 --no-cpu-baseline"
echo "== c2 overlap=1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j_c2.json 2> gpurun_out/r02j_c2.err; echo "rc=$?"
echo "== c2 overlap=0"; timeout 600 python bench.py --steps 10 --warmup 3 --opt overlap_tail=0 $B --no-verify > gpurun_out/r02j_c2_ov0.json 2> gpurun_out/r02j_c2_ov0.err; echo "rc=$?"
echo "== c1"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 $B > gpurun_out/r02j_c1.json 2> gpurun_out/r02j_c1.err; echo "rc=$?"
echo "== c1 overlap=0"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --opt overlap_tail=0 $B > gpurun_out/r02j_c1_ov0.json 2> gpurun_out/r02j_c1_ov0.err; echo "rc=$?"
for f in c2 c2_ov0 c1 c1_ov0; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02j_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'))
    if d.get('finder_api'): print('api', d['finder_api'])
except Exception as e:
    print('no line', e)
PY
done
tail -n 3 gpurun_out/r02j_*.err
