#!/bin/bash
# session 4, call 24 (final): the whole GPU suite, the default bench line, the reference arm, the ncu launch list, smoke()
cd /root/repo
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r03q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r03q_pytest.log
(time python bench.py) > gpurun_out/r03q_bench.json 2> gpurun_out/r03q_bench.err; echo "bench rc=$?"; tail -n 4 gpurun_out/r03q_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r03q_bench.json').read().strip().splitlines()[-1])
def show(o):
    r=o['roofline']
    print(o['config']['name'], 'ms', round(o['ms_per_step'],2), 'frac', round(r['frac'],4), r['kernel'][:40], 'step', r.get('step',{}).get('frac'), 'e2e', round(o['e2e']['ms_per_step'],1), 'pcm', o.get('e2e_pcm16',{}).get('ms_per_step'), 'fast', o.get('fast_math',{}).get('ms_per_step'), 'cpu', o.get('cpu_baseline',{}).get('value'), 'e2e/cpu', o['e2e']['value']/o['cpu_baseline']['value'] if 'cpu_baseline' in o else None)
show(d)
for k in d['roofline'].get('per_plugin',[]): print('  ', k['plugin'], round(k['mean_launch_ms'],2), round(k['frac'],4))
for o in d.get('other_configs',[]):
    if 'error' in o: print(o)
    else: show(o)
"
(time python bench.py --impl reference --steps 2 --warmup 3) > gpurun_out/r03q_bench_reference.json 2> gpurun_out/r03q_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r03q_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r03q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-survey > gpurun_out/r03q_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -n 4
