CB="python tools/chain_bench.py --steps 2 --warmup 1"
for n in 4096 8192 16384 32768; do
for p in lane coop; do
$CB --chain JuicyPunch,JuicyWidth --clips $n --synth drum --path $p
$CB --chain JuicyPunch --clips $n --synth drum --path $p
$CB --chain JuicyWidth --clips $n --synth drum --path $p
$CB --chain JuicyInfer --clips $n --synth mixed --path $p
done; done
