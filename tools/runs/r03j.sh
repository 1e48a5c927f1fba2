#!/bin/bash
# session 4, call 17: JuicyCohere on the few-streams kernel, stages of its block pre-pass left out one at a time (timing only)
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed --chain JuicyCohere --clips 148"
echo -n "all stages: "; $CB | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"
for m in 2 4 16 63; do
  echo -n "skip mask $m: "; JUICY_BATCH_LIB=$V/libjb_skip$m.so $CB | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"
done
