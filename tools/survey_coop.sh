CB="python tools/chain_bench.py --steps 2 --warmup 1"
$CB --chain JuicyInfer --clips 65536 --synth mixed --path coop
$CB --chain JuicyInfer --clips 16384 --synth mixed --path coop
$CB --chain JuicyInfer --clips 4096 --synth mixed --path coop
$CB --chain JuicyWidth --clips 16384 --synth drum --path coop
$CB --chain JuicyWidth --clips 4096 --synth drum --path coop
$CB --chain JuicyPunch --clips 16384 --synth drum --path coop
$CB --chain JuicyPunch,JuicyWidth --clips 32768 --synth drum --path coop
$CB --chain JuicyPunch,JuicyWidth,JuicyInfer --clips 4096 --synth drum --path coop
