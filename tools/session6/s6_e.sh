python -m pytest tests -x -q -m gpu 2>&1 | tail -12
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-40s %6d %-14s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
$CB --chain JuicyTexture --clips 32768 --synth impulse --clipmod material=5 | line "mod5 conc"
JB_GROUP_SERIAL=1 $CB --chain JuicyTexture --clips 32768 --synth impulse --clipmod material=5 | line "mod5 serial"
$CB --chain JuicyTexture --clips 32768 --synth impulse --clipmod material=5 --clipranges | line "ranges conc"
JB_GROUP_SERIAL=1 $CB --chain JuicyTexture --clips 32768 --synth impulse --clipmod material=5 --clipranges | line "ranges serial"
for m in 1 2; do $CB --chain JuicyTexture --clips 6554 --synth impulse --param 0:material=$m | line "6554 m$m"; done
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
$CB --chain $FULL --clips 32768 --synth mixed | line "auto(exact)"
$CB --chain JuicyPunch --clips 65536 --synth drum --math exact | line "exact"
$CB --chain JuicyPunch --clips 65536 --synth drum --math fast | line "fast"
$CB --chain JuicySaturator --clips 65536 --synth sweep --math exact | line "exact"
$CB --chain JuicySaturator --clips 65536 --synth sweep --math fast | line "fast"
$CB --chain JuicyPunch --clips 16384 --synth drum --math exact | line "exact"
$CB --chain JuicySaturator --clips 16384 --synth sweep --math exact | line "exact"
