CB="python tools/chain_bench.py --steps 2 --warmup 1"
$CB --chain JuicyInfer --clips 65536 --synth mixed --path lane
$CB --chain JuicyInfer --clips 16384 --synth mixed --path lane
$CB --chain JuicyWidth --clips 65536 --synth drum --path lane
$CB --chain JuicyWidth --clips 16384 --synth drum --path lane
$CB --chain JuicySaturator --clips 65536 --synth sweep --path lane
$CB --chain JuicySaturator --clips 16384 --synth sweep --path lane
$CB --chain JuicyCohere --clips 65536 --synth noise --path lane
$CB --chain JuicyCohere --clips 16384 --synth noise --path lane
$CB --chain JuicyPunch --clips 16384 --synth drum --path lane
$CB --chain JuicyMotion --clips 16384 --synth drum --path lane
$CB --chain JuicyTexture --clips 8192 --synth impulse --path lane --param 0:material=1
$CB --chain JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer --clips 32768 --synth mixed --path lane
