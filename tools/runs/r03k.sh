#!/bin/bash
# session 4, call 18: few-streams kernel without register spills (5 / 4 CTAs per SM): Saturator, Infer, Cohere, and C1
cd /root/repo
python -m pytest tests/test_gpu_solo.py -m gpu -x -q 2>&1 | tail -n 3
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed"
t() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms [%s]' % (d['ms_per_render'], d['path']))"; }
{
echo -n "C1 Saturator 1 clip x 10 s: "; $CB --chain JuicySaturator --clips 1 --samples 480000 --synth sweep | t
for p in JuicySaturator JuicyInfer JuicyCohere; do for c in 148 592 740 888; do
  echo -n "$p $c clips auto: "; $CB --chain $p --clips $c | t
done; done
for p in JuicySaturator JuicyInfer JuicyCohere; do for c in 740 888; do
  echo -n "$p $c clips lane: "; $CB --chain $p --clips $c --path lane | t
done; done
echo -n "JuicySaturator exact 148: "; $CB --chain JuicySaturator --clips 148 --math exact | t
} | tee gpurun_out/r03k_solo.txt
