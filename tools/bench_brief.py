import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        st=d.get('stage_ms',{})
        print(f, 'ms=%.3f'%d['ms_per_step'], 'e2e=%.2f'%d['e2e']['ms_per_step'], {k:round(v,3) for k,v in st.items() if v>0.001 and k!='total'}, d['result'].get('buckets'))
    except Exception as e: print(f, 'ERR', e)
