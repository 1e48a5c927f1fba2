#!/bin/bash
cd /root/repo
CB="python tools/chain_bench.py --steps 1 --warmup 1 --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum"
for m in exact fast; do
  if [ $m = exact ]; then M=""; else M="--math fast"; fi
  ncu --set full --clock-control none --import-source on -k regex:jb_coop -c 1 -f -o gpurun_out/r02k_$m $CB $M > gpurun_out/r02k_ncu_$m.log 2>&1
  ncu -i gpurun_out/r02k_$m.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r02k_$m.src.csv.gz
  ncu -i gpurun_out/r02k_$m.ncu-rep --page raw --csv > gpurun_out/r02k_$m.raw.csv 2>/dev/null
  ls -la gpurun_out/r02k_$m*
done
