#!/bin/bash
# session 4, call 8: timeline of jb_process_host on the C5 shard (JB_HOST_TRACE), one block per slice and three
cd /root/repo
rm -f gpurun_out/r03a_trace1.txt gpurun_out/r03a_trace3.txt
JB_HOST_TRACE=gpurun_out/r03a_trace1.txt python tools/e2e_sweep.py --chain full --clips 32768 --reps 1 --rounds 6 --pass-mib 32768 --slice-mib 96 | cut -c1-300
JB_HOST_TRACE=gpurun_out/r03a_trace3.txt JB_HOST_MIN_SLICE_BLOCKS=3 python tools/e2e_sweep.py --chain full --clips 32768 --reps 1 --rounds 4 --pass-mib 32768 --slice-mib 96 --extra "JB_HOST_MIN_SLICE_BLOCKS=3" | cut -c1-300
wc -l gpurun_out/r03a_trace*.txt
