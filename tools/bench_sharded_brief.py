import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'n_gpus', d['n_gpus'], 'ms=%.3f' % d['ms_per_step'], 'value=%.3e' % d['value'], 'e2e_ms=%.1f' % d['e2e']['ms_per_step'])
        r = d['phase_ms_per_rank'][0]
        print('   rank0 phases', {k: round(v, 3) for k, v in r.items()})
        print('   rank0 stages', {k: round(v, 3) for k, v in d['stage_ms_rank0_last_job'].items() if v > 0.002})
        print('   nvlink', d['roofline']['nvlink'], d['result'].get('received'))
    except Exception as e:
        print(f, 'ERR', e)
