#!/bin/bash
# session 4, call 21: JuicyCohere few-streams kernel, everything plugin-specific skipped step by step
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
t() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"; }
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed --chain JuicyCohere --clips 148"
for v in skip255 skip511 skip1023; do echo -n "$v: "; JUICY_BATCH_LIB=$V/libjb_$v.so $CB | t; done
