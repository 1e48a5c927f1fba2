#!/bin/bash
# session 4, call 26: few-streams kernel against the lane kernels beyond one wave of CTAs (where to put JB_SOLO_LIMIT)
cd /root/repo
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed"
t() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f' % d['ms_per_render'])"; }
for p in JuicySaturator JuicyInfer JuicyCohere; do for c in 740 888 1184 1480 2072; do
  s=$(JB_SOLO_LIMIT=4000 $CB --chain $p --clips $c | t); l=$($CB --chain $p --clips $c --path lane | t)
  echo "$p $c clips: solo $s ms, lane / pair kernels $l ms"
done; done | tee gpurun_out/r03r_solo_limit.txt
