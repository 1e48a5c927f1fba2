#!/bin/bash
# session 4, call 23: the onset walk from `next` on; cycle counts of all three warp roles; few-streams timings
cd /root/repo
python -m pytest tests/test_gpu_solo.py -m gpu -x -q 2>&1 | tail -n 3
V=juicy-audio-plugins_b200/build/variants
for p in JuicyCohere JuicySaturator; do echo "== $p"; JUICY_BATCH_LIB=$V/libjb_clocks.so python tools/chain_bench.py --steps 1 --warmup 0 --synth mixed --chain $p --clips 148 | grep -v "^{" | cut -c1-400; done | tee gpurun_out/r03p_clocks.txt
CB="python tools/chain_bench.py --steps 3 --warmup 1"
t() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"; }
{
echo -n "C1 Saturator 1 clip x 10 s: "; $CB --chain JuicySaturator --clips 1 --samples 480000 --synth sweep | t
for p in JuicySaturator JuicyInfer JuicyCohere; do for c in 148 592; do echo -n "$p $c clips: "; $CB --synth mixed --chain $p --clips $c | t; done; done
} | tee gpurun_out/r03p_solo.txt
