#!/bin/bash
# round 2, GPU call L (1 GPU): compute-sanitizer (memcheck, racecheck) over the kernels added this round; the driver's own
# bench call (defaults) and the reference arm; dup with parity
set -u
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/r02l_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r02l_plain.log
echo "== memcheck"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02l_memcheck.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02l_memcheck.log
echo "== racecheck"; timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02l_racecheck.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02l_racecheck.log
echo "== bench (driver call)"; timeout 900 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "rc=$?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err; echo "rc=$?"
echo "== dup --verify"; timeout 900 python bench.py --config dup --verify --steps 5 --warmup 3 --api-parts 0 --no-cpu-baseline > gpurun_out/r02l_dup.json 2> gpurun_out/r02l_dup.err; echo "rc=$?"
for f in bench dup; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02l_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), 'parity', d.get('parity',{}).get('vs_oracle'), 'launches', d['gpu_launches'], 'clocks', d.get('clocks'))
except Exception as e:
    print('no line', e)
PY
done
tail -n 3 gpurun_out/r02l_*.err
