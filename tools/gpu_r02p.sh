#!/bin/bash
# round 2, GPU call P (1 GPU): second version of the warp-private filter (record pairs, check-free batches, per-lane suspect masks):
# parity, sanitizer, c2 A/B against the first version and with smaller leaves, ncu
set -u
TAG=${1:-r02p}
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_fw.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_memcheck.log
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_fw.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_racecheck.log
B="--api-parts 0 --no-cpu-baseline"
run() { name=$1; shift; echo "== $name"; timeout 600 python bench.py --steps 10 --warmup 3 $B "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "rc=$?"; }
run c2_fw2 --opt filter_warp=2
run c2_fw1 --opt filter_warp=1
run c2_fw2_t400 --opt filter_warp=2 --opt warp_leaf_target=400
run c2_fw1_t400 --opt filter_warp=1 --opt warp_leaf_target=400
run c1_fw2 --config c1
run c4_fw2 --config c4 --steps 5
run c3_fw2 --config c3 --steps 5
for f in c2_fw2 c2_fw1 c2_fw2_t400 c2_fw1_t400 c1_fw2 c4_fw2 c3_fw2; do echo "-- $f"; python - <<PY
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
