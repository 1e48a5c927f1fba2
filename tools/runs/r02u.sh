#!/bin/bash
# session 4, call 2: new jb_process_host geometry (single pass, >= 3 blocks per slice, tapered tail) + the C5-first bench line
cd /root/repo
X='JB_HOST_MIN_SLICE_BLOCKS=2,JB_HOST_TAPER=0|JB_HOST_MIN_SLICE_BLOCKS=2,JB_HOST_TAPER=1|JB_HOST_MIN_SLICE_BLOCKS=3,JB_HOST_TAPER=0|JB_HOST_MIN_SLICE_BLOCKS=3,JB_HOST_TAPER=1|JB_HOST_MIN_SLICE_BLOCKS=4,JB_HOST_TAPER=1|JB_HOST_MIN_SLICE_BLOCKS=6,JB_HOST_TAPER=1'
python tools/e2e_sweep.py --chain full --clips 32768 --floor --pass-mib 32768 --slice-mib 96 --extra "$X" > gpurun_out/r02u_c5.txt 2> gpurun_out/r02u_c5.err; echo "c5 rc=$?"
X2='JB_HOST_TAPER=0|JB_HOST_TAPER=1'
python tools/e2e_sweep.py --chain JuicyPunch,JuicyWidth --synth drum --clips 4096 --floor --reps 5 --pass-mib 32768 --slice-mib 64,96,128 --extra "$X2" > gpurun_out/r02u_c2.txt 2> gpurun_out/r02u_c2.err; echo "c2 rc=$?"
cat gpurun_out/r02u_c5.txt gpurun_out/r02u_c2.txt; tail -n 3 gpurun_out/r02u_c5.err gpurun_out/r02u_c2.err
(time python bench.py --steps 5 --no-cpu --no-survey) > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/r02u_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[-1])
print(d['config']['name'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'pcm', d.get('e2e_pcm16',{}).get('ms_per_step'), 'fast', d.get('fast_math',{}).get('ms_per_step'))
r=d['roofline']; print(r['kernel'], r['frac'], r['mean_launch_ms'], r.get('step'))
for k in r.get('per_plugin',[]): print(k['plugin'], round(k['mean_launch_ms'],2), round(k['frac'],4))
"
