import json, sys
d = json.loads(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.3e samples/s  %.3f ms/step  e2e %.3e  launches/step %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d.get('launches_per_step')))
for k, v in d['kernels'].items():
    r = d.get('roofline_all', {}).get(k)
    print('%-18s %3d x %8.1f us = %7.3f ms %s' % (k, v['launches_per_step'], v['us_per_launch'], v['ms_per_step'],
                                                 ('%5.1f%% of %s peak' % (100 * r['frac'], r['bound'])) if r else ''))
if d.get('fastgen'):
    print({k: v for k, v in d['fastgen'].items() if k != 'note'})
if d.get('cpu_baseline'):
    print('cpu', d['cpu_baseline']['value'])
if d.get('batch4'):
    print('batch4', d['batch4'])
