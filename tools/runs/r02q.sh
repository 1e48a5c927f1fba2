#!/bin/bash
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
fmt() { grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-44s %8.3f ms %5.1f%% [%s]' % (sys.argv[1], d['ms_per_render'], 100*d['frac_of_measured_hbm'], d['path']))
" "$1"; }
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed --inplace"
{
for lib in t1 t0; do
  L="JUICY_BATCH_LIB=$V/libjb_$lib.so"
  env $L $CB --chain JuicyInfer --clips 65536 2>&1 | fmt "$lib Infer 65536"
  for p in JuicySaturator JuicyCohere JuicyWidth JuicyPunch; do
    env $L JB_TILE=1 $CB --chain $p --clips 65536 2>&1 | fmt "$lib $p 65536 JB_TILE=1"
  done
  for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer; do
    env $L JB_TILE=1 $CB --chain $p --clips 32768 2>&1 | fmt "$lib $p 32768 JB_TILE=1"
  done
  for p in JuicySaturator JuicyCohere JuicyWidth; do
    env $L JB_TILE=1 $CB --chain $p --clips 16384 2>&1 | fmt "$lib $p 16384 JB_TILE=1"
  done
done
L="JUICY_BATCH_LIB=$V/libjb_t1.so"
for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer; do
  env $L $CB --chain $p --clips 32768 2>&1 | fmt "default path $p 32768"
done
for p in JuicySaturator JuicyCohere JuicyWidth; do
  env $L $CB --chain $p --clips 16384 2>&1 | fmt "default path $p 16384"
done
} | tee gpurun_out/r02q2_bench.txt
