#!/bin/bash
# round 2, GPU call S (1 GPU): early fetch by an SM-side copy kernel (A/B), the warp filter with room for 192 suspects on config 3
# (A/B against the CTA-per-leaf filter), cooperative grouping above 2^20 candidates, smaller edge table: parity, A/B
set -u
TAG=${1:-r02s}
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py tests/test_gpu_finder.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
B="--api-parts 0 --no-cpu-baseline"
run() { name=$1; shift; echo "== $name"; timeout 600 python bench.py --steps 10 --warmup 3 $B "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "rc=$?"; }
run c2
run c2_ef0 --opt early_fetch=0
run c2_b
run c2_ef0_b --opt early_fetch=0
run c3 --config c3 --steps 5 --verify
run c3_fw2 --config c3 --steps 5 --opt filter_warp=2
run c3_coop --config c3 --steps 5 --opt coop_group=2
run c1 --config c1
run c4 --config c4 --steps 5
for f in c2 c2_ef0 c2_b c2_ef0_b c3 c3_fw2 c3_coop c1 c4; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), [round(x,3) for x in d['e2e']['ms_load_build_fetch']], 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'res', {k:d['result'].get(k) for k in ('buckets','oversized','candidates')})
except Exception as e:
    print('no line', e)
PY
done
tail -n 3 gpurun_out/${TAG}_*.err
