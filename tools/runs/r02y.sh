#!/bin/bash
# session 4, call 6: the default bench line with the median e2e and the same-run PCIe floor
cd /root/repo
(time python bench.py) > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?"; tail -n 4 gpurun_out/r02y_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1])
def show(o):
    r=o['roofline']; e=o['e2e']
    print(o['config']['name'], 'ms', round(o['ms_per_step'],2), 'frac', round(r['frac'],4), 'traffic', r.get('traffic'), 'step', r.get('step',{}).get('frac'), 'e2e med/mean/min', round(e['ms_per_step'],1), round(e['ms_per_step_mean'],1), round(e['ms_per_step_min'],1), 'floor', e.get('pcie_floor_ms'), 'pcm', o.get('e2e_pcm16',{}).get('ms_per_step'), 'fast', o.get('fast_math',{}).get('ms_per_step'), 'e2e/cpu', o['e2e']['value']/o['cpu_baseline']['value'] if 'cpu_baseline' in o else None)
show(d)
for o in d.get('other_configs',[]):
    if 'error' in o: print(o)
    else: show(o)
"
