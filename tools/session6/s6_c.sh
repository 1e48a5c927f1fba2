python -m pytest tests -x -q -m gpu 2>&1 | tail -8
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-40s %6d %-10s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for m in 2 3; do
  $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=$m | line "single m$m"
done
$CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 | line "C3 mod5"
$CB --chain JuicyTexture --clips 32768 --synth impulse --clipmod material=5 | line "mod5"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
$CB --chain $FULL --clips 32768 --synth mixed | line split
$CB --chain $FULL --clips 4096 --synth mixed | line split
python bench.py --steps 5 --warmup 3
