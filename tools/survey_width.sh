CB="python tools/chain_bench.py --steps 2 --warmup 1"
$CB --chain JuicyWidth --clips 65536 --synth drum --path lane
$CB --chain JuicyWidth --clips 16384 --synth drum --path lane
$CB --chain JuicyPunch,JuicyWidth --clips 32768 --synth drum --path lane
