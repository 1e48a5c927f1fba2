#!/bin/bash
# session 4, call 7: bench.py's e2e (370 ms at a 280 ms floor) against tools/e2e_sweep.py's (265 ms): nvidia-smi poll / CPU binding, alternated
cd /root/repo
S="python tools/e2e_sweep.py --chain full --clips 32768 --reps 1 --rounds 5 --pass-mib 32768 --slice-mib 96 --floor"
for r in 1 2 3; do
echo "plain";        $S | cut -c1-400
echo "smi 100 ms";   $S --smi-ms 100 | cut -c1-400
echo "bind";         $S --bind | cut -c1-400
echo "smi 1000 ms";  $S --smi-ms 1000 | cut -c1-400
done 2>&1 | grep -v wall_s | tee gpurun_out/r02z_e2e.txt
