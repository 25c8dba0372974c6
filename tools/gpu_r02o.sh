#!/bin/bash
# round 2, GPU call O (1 GPU): warp-private filter with pipelined loads and ten-warp CTAs: parity, c2 A/B, ncu
set -u
TAG=${1:-r02o}
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
B="--api-parts 0 --no-cpu-baseline"
for v in 1 0; do
  echo "== c2 filter_warp=$v"; timeout 600 python bench.py --steps 10 --warmup 3 --opt filter_warp=$v $B > gpurun_out/${TAG}_c2_fw$v.json 2> gpurun_out/${TAG}_c2_fw$v.err; echo "rc=$?"
done
echo "== c1"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 $B > gpurun_out/${TAG}_c1_fw1.json 2> gpurun_out/${TAG}_c1_fw1.err; echo "rc=$?"
for f in c2_fw1 c2_fw0 c1_fw1; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'res', {k:d['result'].get(k) for k in ('buckets','oversized','filter_warp','candidates')})
except Exception as e:
    print('no line', e)
PY
done
CMD="python tools/profile_target.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_filter_warp' -c 2 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
tail -n 3 gpurun_out/${TAG}_*.err
