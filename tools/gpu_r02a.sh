#!/bin/bash
# round 2, GPU call A (1 GPU): parity tests, full-size verification, A/B of the part-aligned extraction kernel, ncu capture
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log
echo "== bench c2 (new kernel)"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; echo "rc=$?"
python tools/bench_brief.py gpurun_out/r02a_bench_c2.json 2>/dev/null | head -20
echo "== bench c2 (byte-position kernel)"; timeout 600 python bench.py --steps 10 --warmup 3 --opt uniform_kernel=0 --no-verify --api-parts 0 --no-cpu-baseline > gpurun_out/r02a_bench_c2_old.json 2> gpurun_out/r02a_bench_c2_old.err; echo "rc=$?"
echo "== bench c1"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --api-parts 0 --no-cpu-baseline > gpurun_out/r02a_bench_c1.json 2> gpurun_out/r02a_bench_c1.err; echo "rc=$?"
echo "== bench c3 --verify"; timeout 900 python bench.py --config c3 --verify --steps 10 --warmup 3 --api-parts 0 --no-cpu-baseline > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err; echo "rc=$?"
echo "== bench dup --verify"; timeout 900 python bench.py --config dup --verify --steps 5 --warmup 3 --api-parts 0 --no-cpu-baseline > gpurun_out/r02a_bench_dup.json 2> gpurun_out/r02a_bench_dup.err; echo "rc=$?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; echo "rc=$?"
echo "== ncu"; timeout 900 bash tools/ncu_c2.sh r02a > gpurun_out/r02a_ncu.log 2>&1; echo "rc=$?"
echo "== bench c4 1 GPU --verify"; timeout 1500 python bench.py --config c4 --verify --steps 5 --warmup 3 --api-parts 0 --no-cpu-baseline > gpurun_out/r02a_bench_c4.json 2> gpurun_out/r02a_bench_c4.err; echo "rc=$?"
for f in c2 c2_old c1 c3 dup c4; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02a_bench_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'parity', d.get('parity',{}).get('vs_oracle'), d.get('parity',{}).get('oracle'), 'uniform', d['result'].get('uniform_kernel'))
except Exception as e:
    print('no line', e)
PY
done
tail -3 gpurun_out/r02a_bench_*.err
