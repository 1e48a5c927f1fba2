python tools/dbg_wood.py
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-40s %6d %-10s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
$CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 | line "mod5 conc"
JB_GROUP_SERIAL=1 $CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 | line "mod5 serial"
$CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 --clipranges | line "ranges conc"
JB_GROUP_SERIAL=1 $CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 --clipranges | line "ranges serial"
for m in 0 1 2 3 4; do
$CB --chain JuicyTexture --clips 1638 --synth impulse --param 0:material=$m | line "1638 m$m"
done
