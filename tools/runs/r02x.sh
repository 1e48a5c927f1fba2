#!/bin/bash
# session 4, call 5: what separates bench.py's e2e (374 ms) from tools/e2e_sweep.py's (265 ms) on the same geometry -- the
# nvidia-smi poll, the CPU binding, or the box; ncu --set full of the C5 shard's Punch launch
cd /root/repo
S="python tools/e2e_sweep.py --chain full --clips 32768 --reps 1 --rounds 6 --pass-mib 32768 --slice-mib 96"
for r in 1 2; do
echo "plain";        $S | cut -c1-330
echo "smi 100 ms";   $S --smi-ms 100 | cut -c1-330
echo "bind";         $S --bind | cut -c1-330
echo "bind + smi";   $S --bind --smi-ms 100 | cut -c1-330
done 2>&1 | tee gpurun_out/r02x_e2e.txt
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
ncu --set full --clock-control none --import-source on -k jb_pair_kernel -s 1 -c 1 -f -o gpurun_out/r02x_punch_pair python tools/chain_bench.py --steps 1 --warmup 1 --chain $FULL --clips 32768 --synth mixed --inplace > gpurun_out/r02x_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/r02x_punch_pair.ncu-rep gpurun_out/r02x_punch_pair_ncu.json; echo "summary rc=$?"
rm -f gpurun_out/r02x_punch_pair.ncu-rep.tmp; ls -la gpurun_out/r02x*
