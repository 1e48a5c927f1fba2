#!/bin/bash
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
fmt() { grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-40s %8.3f ms [%s]' % (sys.argv[1], d['ms_per_render'], d['path']))
" "$1"; }
CB="python tools/chain_bench.py --steps 5 --warmup 2 --clips 4096 --chain JuicyPunch,JuicyWidth"
{
for lib in h8 h4 h2 h1; do
  L="JUICY_BATCH_LIB=$V/libjb_$lib.so"
  for iso in 0 1; do
    env $L JB_CO_ISOLATE=$iso $CB --synth drum 2>&1 | fmt "$lib exact drum iso=$iso"
  done
  env $L $CB --synth mixed 2>&1 | fmt "$lib exact mixed"
done
} | tee gpurun_out/r02p_bench.txt
