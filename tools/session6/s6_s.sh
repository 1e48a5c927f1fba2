CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane --inplace"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
for seg in 4 8 16 32 47; do
  for c in 4096 16384; do
    JB_PIPE_SEGMENTS=$seg $CB --chain $FULL --clips $c --synth mixed | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('segments', $seg, 'clips', d['clips'], '%.2f ms' % d['ms_per_render'])"
  done
done
