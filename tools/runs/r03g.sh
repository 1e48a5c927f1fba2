#!/bin/bash
# session 4, call 14: the default bench line (final form: 1 Hz nvidia-smi poll during the end-to-end steps) and the reference arm
cd /root/repo
(time python bench.py) > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err; echo "bench rc=$?"; tail -n 4 gpurun_out/r03g_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r03g_bench.json').read().strip().splitlines()[-1])
def show(o):
    r=o['roofline']; e=o['e2e']
    print(o['config']['name'], 'ms', round(o['ms_per_step'],2), 'frac', round(r['frac'],4), 'step', r.get('step',{}).get('frac'), 'e2e med/mean/min/floor', round(e['ms_per_step'],1), round(e['ms_per_step_mean'],1), round(e['ms_per_step_min'],1), round(e['pcie_floor_ms'],1), 'pcm', o.get('e2e_pcm16',{}).get('ms_per_step'), 'fast', o.get('fast_math',{}).get('ms_per_step'), 'e2e/cpu', o['e2e']['value']/o['cpu_baseline']['value'] if 'cpu_baseline' in o else None, o['clocks'])
show(d)
for o in d.get('other_configs',[]):
    if 'error' in o: print(o)
    else: show(o)
"
(time python bench.py --impl reference --steps 2 --warmup 3) > gpurun_out/r03g_bench_reference.json 2> gpurun_out/r03g_bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r03g_bench_reference.json
