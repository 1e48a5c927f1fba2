#!/bin/bash
cd /root/repo
CB="python tools/chain_bench.py --steps 1 --warmup 1 --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum"
for m in exact; do
  ncu --set full --clock-control none --import-source on -k regex:jb_coop -c 1 -f -o gpurun_out/r02m_$m $CB > gpurun_out/r02m_ncu_$m.log 2>&1
  ncu -i gpurun_out/r02m_$m.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r02m_$m.src.csv.gz
  ncu -i gpurun_out/r02m_$m.ncu-rep --page raw --csv > gpurun_out/r02m_$m.raw.csv 2>/dev/null
done
rm -f gpurun_out/r02m_exact.ncu-rep.tmp
