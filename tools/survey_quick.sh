CB="python tools/chain_bench.py --steps 2 --warmup 1"
$CB --chain JuicySaturator --clips 65536 --synth sweep --path lane
$CB --chain JuicySaturator --clips 16384 --synth sweep --path lane
$CB --chain JuicyPunch --clips 16384 --synth drum --path lane
$CB --chain JuicyPunch --clips 65536 --synth drum --path lane
$CB --chain JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer --clips 32768 --synth mixed --path lane
