#!/bin/bash
# round 2, GPU call T (1 GPU): level 2 with the digit taken from the record in the flush (A/B against a build with the digit array),
# the filter file reduced to its final version, config 3 / 4 without the early fetch: parity, A/B
set -u
TAG=${1:-r02t}
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
B="--api-parts 0 --no-cpu-baseline"
run() { name=$1; shift; echo "== $name"; timeout 600 python bench.py --steps 10 --warmup 3 $B "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "rc=$?"; }
run c2
NRPCALC_B200_LIB=$PWD/_variants/libnrpb200_keydig0.so run c2_keydig0
run c2_b
NRPCALC_B200_LIB=$PWD/_variants/libnrpb200_keydig0.so run c2_keydig0_b
run c3 --config c3 --steps 5
run c4 --config c4 --steps 5
NRPCALC_B200_LIB=$PWD/_variants/libnrpb200_keydig0.so run c4_keydig0 --config c4 --steps 5
for f in c2 c2_keydig0 c2_b c2_keydig0_b c3 c4 c4_keydig0; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), [round(x,3) for x in d['e2e']['ms_load_build_fetch']], 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'roof', round(d['roofline']['frac'],3), d['roofline']['kernel'])
except Exception as e:
    print('no line', e)
PY
done
tail -n 3 gpurun_out/${TAG}_*.err
