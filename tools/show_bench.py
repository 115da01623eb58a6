"""Print the headline fields and the per-kernel table of bench.py JSON lines:  python tools/show_bench.py FILE..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    r = d.get('roofline') or {}
    print(f, 'value', round(d['value']), d['unit'], 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
          'roofline', r.get('kernel'), round(r.get('frac', 0), 3), 'n_gpus', d['n_gpus'])
    for k, v in (d.get('kernels') or {}).items():
        print('   %-20s %.3f ms x%.0f' % (k, v['ms_per_step'], v['launches_per_step']))
