#!/bin/bash
# round 2, GPU call V (1 GPU): what the driver runs at round end, with the final code: whole -m gpu suite, smoke(), bench.py with its
# defaults; then a full capture of the three bulk kernels of one config-2 build and its ncu launch list
set -u
TAG=${1:-r02v}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 330 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
echo "== bench (driver call)"; timeout 240 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'value', d['value'], 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), d['roofline']['kernel'], 'pipeline', round(d['roofline']['pipeline']['frac'],3), 'parity', d.get('parity',{}).get('vs_oracle'), 'launches', d['gpu_launches'], 'clocks', d.get('clocks'))
    print('stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0})
    print('cpu', d.get('cpu_baseline'), 'api', d.get('finder_api'))
except Exception as e:
    print('no line', e)
PY
CMD="python tools/profile_target.py"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_filter_warp|k_split_from_records|k_split_uniform' -c 3 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$?"
tail -n 3 gpurun_out/${TAG}_*.err
