# end-of-session survey: every plugin alone at 65536 / 8192 clips, C1..C5, through tools/chain_bench.py (device resident)
CB="python tools/chain_bench.py --steps 2 --warmup 1"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-44s %6d %-18s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer JuicyPunch JuicyMotion JuicyTexture; do
  $CB --chain $p --clips 65536 --synth mixed | line "65536"
  $CB --chain $p --clips 8192 --synth mixed | line "8192"
done
$CB --chain JuicySaturator --clips 1 --samples 480000 --synth sweep | line "C1"
$CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum | line "C2 (auto: coop)"
$CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum --path lane | line "C2 (lane kernels)"
$CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 | line "C3 mod5"
for m in 0 1 2 3 4; do $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=$m | line "C3 material $m"; done
$CB --chain JuicyInfer --clips 65536 --synth mixed --inplace | line "C4 in place"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
$CB --chain $FULL --clips 32768 --synth mixed --inplace | line "C5 shard (auto=fast)"
$CB --chain $FULL --clips 32768 --synth mixed --inplace --math exact | line "C5 shard exact"
$CB --chain $FULL --clips 32768 --synth mixed --inplace --param 2:material=2 | line "C5 shard wood (auto=exact)"
$CB --chain $FULL --clips 4096 --synth mixed --inplace | line "chain 4096"
