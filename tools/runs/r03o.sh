#!/bin/bash
# session 4, call 22: cycle counts inside the few-streams kernel (helpers' stages, the envelope walker's waits)
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
for p in JuicyCohere JuicySaturator JuicyInfer; do echo "== $p"; JUICY_BATCH_LIB=$V/libjb_clocks.so python tools/chain_bench.py --steps 1 --warmup 0 --synth mixed --chain $p --clips 148 | cut -c1-400; done | tee gpurun_out/r03o_clocks.txt
