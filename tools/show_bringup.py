import json, sys
r = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/learner_bringup.json'))
for k, v in r.items():
    if k.endswith('_sec'):
        continue
    if isinstance(v, dict) and 'error' in v:
        print(k, 'ERROR\n', v['error']); continue
    if k.startswith('compute'):
        print(k, ' '.join('%s %.1e' % (q.replace('grad/', 'g/'), e['rel']) for q, e in v.items()))
    elif k.startswith('schedule'):
        print(k)
        for rec in v:
            print('  u%d gs %d->%d params %.1e step %.1e' % (rec['update'], rec['gs_before'], rec['gs_after'], rec['params_rel'], rec['step_rel']),
                  'clip', rec.get('clip_coeff'), 'gn', rec.get('grad_norm'))
            for kk in ('precon', 'inv_A', 'inv_G', 'sums_A', 'sums_G'):
                if kk in rec:
                    print('       ', kk, ' '.join('%s %.1e' % (a, b) for a, b in rec[kk].items()))
    else:
        print(k, v)
