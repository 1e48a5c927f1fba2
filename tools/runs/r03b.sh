#!/bin/bash
# session 4, call 9: timeline of jb_process_host_pcm16 on the C5 shard, one block per slice and three
cd /root/repo
rm -f gpurun_out/r03b_trace1.txt gpurun_out/r03b_trace3.txt
JB_HOST_TRACE=gpurun_out/r03b_trace1.txt python tools/e2e_sweep.py --pcm16 --chain full --clips 32768 --reps 1 --rounds 3 --pass-mib 32768 --slice-mib 96 | cut -c1-300
JB_HOST_TRACE=gpurun_out/r03b_trace3.txt python tools/e2e_sweep.py --pcm16 --chain full --clips 32768 --reps 1 --rounds 3 --pass-mib 32768 --slice-mib 96 --extra "JB_HOST_MIN_SLICE_BLOCKS=3" | cut -c1-300
nsys --version 2>/dev/null | head -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pcm16 -c 12 python tools/e2e_sweep.py --pcm16 --chain JuicyInfer --clips 32768 --reps 1 --rounds 1 --pass-mib 32768 --slice-mib 96 2>&1 | grep -E "jb_|gpu__time" | head -30
