#!/bin/bash
# round 2, GPU call Q (1 GPU): third version of the warp-private filter (leaf bits in registers) and the split flush with eight
# records per thread in flight (A/B against a build without it): parity, c2 A/B, ncu
set -u
TAG=${1:-r02q}
mkdir -p gpurun_out
echo "== pytest (filter_warp default)"; timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
B="--api-parts 0 --no-cpu-baseline"
run() { name=$1; shift; echo "== $name"; timeout 600 python bench.py --steps 10 --warmup 3 $B "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "rc=$?"; }
run c2_fw3 --opt filter_warp=3
run c2_fw2 --opt filter_warp=2
NRPCALC_B200_LIB=$PWD/_variants/libnrpb200_flush0.so run c2_fw3_flush0 --opt filter_warp=3
NRPCALC_B200_LIB=$PWD/_variants/libnrpb200_flush0.so run c2_fw2_flush0 --opt filter_warp=2
run c2_fw3_b --opt filter_warp=3
run c1_fw3 --config c1 --opt filter_warp=3
run c4_fw3 --config c4 --steps 5 --opt filter_warp=3
run c3_fw3 --config c3 --steps 5 --opt filter_warp=3
for f in c2_fw3 c2_fw2 c2_fw3_flush0 c2_fw2_flush0 c2_fw3_b c1_fw3 c4_fw3 c3_fw3; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'res', {k:d['result'].get(k) for k in ('buckets','oversized','candidates')})
except Exception as e:
    print('no line', e)
PY
done
CMD="python tools/profile_target.py"
NRP_PROFILE_OPTS="filter_warp=3" timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_filter_warp|k_split_from_records|k_split_uniform' -c 6 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
tail -n 3 gpurun_out/${TAG}_*.err
