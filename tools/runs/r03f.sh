#!/bin/bash
# session 4, call 13: bench.py's e2e with and without the nvidia-smi poll during the e2e steps, alternated
cd /root/repo
for r in 1 2; do for v in 1 0; do
JB_BENCH_SMI_E2E=$v python bench.py --steps 7 --no-cpu --no-survey 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; p=d['e2e_pcm16']
print('smi during e2e = $v | e2e median %.1f mean %.1f min %.1f max %.1f floor %.1f | pcm16 median %.1f min %.1f' % (e['ms_per_step'], e['ms_per_step_mean'], e['ms_per_step_min'], e['ms_per_step_max'], e['pcie_floor_ms'], p['ms_per_step'], p['ms_per_step_min']))
"
done; done | tee gpurun_out/r03f_smi.txt
