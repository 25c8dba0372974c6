#!/bin/bash
# round 2, GPU call I (1 GPU): folded flags with the blocked-Bloom probe and deferred confirmation, general path (any alphabet, any k), packed entry;
# whole GPU suite, c3 A/B with full-size parity, c2 / c1, launch list + full capture of the folded kernel on config 3
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02i_pytest.log
B="--api-parts 0 --no-cpu-baseline"
echo "== c3 fold=1 --verify"; timeout 900 python bench.py --config c3 --verify --steps 10 --warmup 3 $B > gpurun_out/r02i_c3_fold1.json 2> gpurun_out/r02i_c3_fold1.err; echo "rc=$?"
echo "== c3 fold=0"; timeout 600 python bench.py --config c3 --no-verify --steps 10 --warmup 3 --opt fold_flags=0 $B > gpurun_out/r02i_c3_fold0.json 2> gpurun_out/r02i_c3_fold0.err; echo "rc=$?"
echo "== c2"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02i_c2.json 2> gpurun_out/r02i_c2.err; echo "rc=$?"
echo "== c1"; timeout 300 python bench.py --config c1 --steps 20 --warmup 3 $B > gpurun_out/r02i_c1.json 2> gpurun_out/r02i_c1.err; echo "rc=$?"
for f in c3_fold1 c3_fold0 c2 c1; do echo "-- $f"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02i_$f.json').read().strip().splitlines()[-1])
    print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'stage', {k:round(v,4) for k,v in d['stage_ms'].items() if v>0}, 'parity', d.get('parity',{}).get('vs_oracle'), 'uniform', d['result'].get('uniform_kernel'))
except Exception as e:
    print('no line', e)
PY
done
echo "== ncu c3"
export NRP_PROFILE_PARTS=1000000 NRP_PROFILE_K=16 NRP_PROFILE_LEN=200 NRP_PROFILE_BG=5000000
CMD="python tools/profile_target.py"
$CMD > gpurun_out/r02i_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02i_plain.log; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02i_launches_c3.csv $CMD > gpurun_out/r02i_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_split_uniform' -c 2 -o gpurun_out/prof_r02i -f $CMD > gpurun_out/r02i_ncu_full.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/r02i_*.err
