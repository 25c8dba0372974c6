#!/bin/bash
# round 2, GPU call N (1 GPU): warp-private duplicate filter (filter_warp.cuh): parity suite, c2 / c1 A/B, ncu capture
set -u
mkdir -p gpurun_out
echo "== pytest (build, fullsize, finder, general)"; timeout 1500 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py tests/test_gpu_finder.py tests/test_gpu_general.py -q -m gpu > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02n_pytest.log
B="--api-parts 0 --no-cpu-baseline"
for v in 1 0; do
  echo "== c2 filter_warp=$v"; timeout 600 python bench.py --steps 10 --warmup 3 --opt filter_warp=$v $B > gpurun_out/r02n_c2_fw$v.json 2> gpurun_out/r02n_c2_fw$v.err; echo "rc=$?"
  echo "== c1 filter_warp=$v"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --opt filter_warp=$v $B > gpurun_out/r02n_c1_fw$v.json 2> gpurun_out/r02n_c1_fw$v.err; echo "rc=$?"
done
for f in c2_fw1 c2_fw0 c1_fw1 c1_fw0; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'res', {k:d['result'].get(k) for k in ('buckets','oversized','filter_warp','candidates')})
except Exception as e:
    print('no line', e)
PY
done
echo "== ncu c2"
CMD="python tools/profile_target.py"
$CMD > gpurun_out/r02n_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02n_plain.log; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r02n.csv $CMD > gpurun_out/r02n_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_filter_warp|k_split_from_records|k_split_uniform' -c 6 -o gpurun_out/prof_r02n -f $CMD > gpurun_out/r02n_ncu_full.log 2>&1; echo "full rc=$?"
tail -n 3 gpurun_out/r02n_*.err
