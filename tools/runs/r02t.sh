#!/bin/bash
# session 4, call 1: jb_process_host slice / pass geometry on the C5 shard and on C2
cd /root/repo
python tools/e2e_sweep.py --chain full --clips 32768 --floor --pass-mib 8192,16384 --slice-mib 96,192,384,768,1536 > gpurun_out/r02t_c5.txt 2> gpurun_out/r02t_c5.err; echo "c5 rc=$?"
python tools/e2e_sweep.py --chain JuicyPunch,JuicyWidth --synth drum --clips 4096 --floor --reps 5 --pass-mib 8192 --slice-mib 16,32,64,96,192,384 > gpurun_out/r02t_c2.txt 2> gpurun_out/r02t_c2.err; echo "c2 rc=$?"
cat gpurun_out/r02t_c5.txt gpurun_out/r02t_c2.txt; tail -3 gpurun_out/r02t_c5.err gpurun_out/r02t_c2.err
