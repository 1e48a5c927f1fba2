CB="python tools/chain_bench.py --steps 2 --warmup 1"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
for n in 4096 32768; do
JB_LANE_SPLIT=0 $CB --chain $FULL --clips $n --synth mixed --path lane
JB_LANE_SPLIT=1 $CB --chain $FULL --clips $n --synth mixed --path lane
done
JB_LANE_SPLIT=0 $CB --chain JuicySaturator,JuicyCohere --clips 65536 --synth noise --path lane
JB_LANE_SPLIT=1 $CB --chain JuicySaturator,JuicyCohere --clips 65536 --synth noise --path lane
JB_LANE_SPLIT=0 $CB --chain JuicySaturator,JuicyWidth,JuicyCohere,JuicyInfer --clips 16384 --synth drum --path lane
JB_LANE_SPLIT=1 $CB --chain JuicySaturator,JuicyWidth,JuicyCohere,JuicyInfer --clips 16384 --synth drum --path lane
