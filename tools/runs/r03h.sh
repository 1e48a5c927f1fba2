#!/bin/bash
# session 4, call 15: the pipeline timeline inside bench.py (why its e2e steps are slower than tools/e2e_sweep.py's)
cd /root/repo
rm -f gpurun_out/r03h_trace.txt
JB_HOST_TRACE=gpurun_out/r03h_trace.txt python bench.py --steps 5 --no-cpu --no-survey 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; p=d['e2e_pcm16']
print('e2e median %.1f mean %.1f min %.1f max %.1f floor %.1f | pcm16 median %.1f min %.1f' % (e['ms_per_step'], e['ms_per_step_mean'], e['ms_per_step_min'], e['ms_per_step_max'], e['pcie_floor_ms'], p['ms_per_step'], p['ms_per_step_min']))
"
wc -l gpurun_out/r03h_trace.txt
