#!/bin/bash
# session 4, call 19: JuicyCohere few-streams kernel with every stage skipped: shared-memory size / the stage barriers
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
t() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"; }
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed --chain JuicyCohere --clips 148"
for v in skip63 small63 skip127; do
  echo -n "$v: "; JUICY_BATCH_LIB=$V/libjb_$v.so $CB | t
done
echo -n "skip63, sweep input: "; JUICY_BATCH_LIB=$V/libjb_skip63.so python tools/chain_bench.py --steps 3 --warmup 1 --synth sweep --chain JuicyCohere --clips 148 | t
echo -n "full, sweep input: "; python tools/chain_bench.py --steps 3 --warmup 1 --synth sweep --chain JuicyCohere --clips 148 | t
echo -n "Saturator, sweep input: "; python tools/chain_bench.py --steps 3 --warmup 1 --synth sweep --chain JuicySaturator --clips 148 | t
