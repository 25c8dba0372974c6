#!/bin/bash
# round 2, GPU call X (1 GPU): the API-level block of the bench line (nrpcalc_b200.finder / finder_packed wall clock on 100k parts) now
# that a fresh process finds the host tail's self-check verdicts remembered
set -u
mkdir -p gpurun_out
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity',{}).get('vs_oracle'))
print('api', d.get('finder_api'))
PY
tail -n 3 gpurun_out/r02x_bench.err
