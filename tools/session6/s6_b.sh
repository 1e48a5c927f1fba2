# single-plugin kernels: parity, then A/B against the generic kernel (JB_LANE_GENERIC=1)
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-40s %6d %-8s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for m in 0 1 2 3 4; do
  $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=$m | line "single m$m"
done
JB_LANE_GENERIC=1 $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=0 | line "generic m0"
$CB --chain JuicyTexture --clips 32768 --synth impulse --param 0:material=0 | line "single m0"
$CB --chain JuicyTexture --clips 32768 --synth impulse --param 0:material=1 | line "single m1"
for c in 16384 32768; do
$CB --chain JuicyMotion --clips $c --synth drum | line single
done
JB_LANE_GENERIC=1 $CB --chain JuicyMotion --clips 16384 --synth drum | line generic
for p in JuicyPunch JuicySaturator JuicyWidth JuicyCohere JuicyInfer; do
  $CB --chain $p --clips 65536 --synth mixed | line single
  JB_LANE_GENERIC=1 $CB --chain $p --clips 65536 --synth mixed | line generic
  $CB --chain $p --clips 16384 --synth mixed | line single
done
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
$CB --chain $FULL --clips 32768 --synth mixed | line split
JB_LANE_SPLIT=1 $CB --chain $FULL --clips 4096 --synth mixed | line split
JB_LANE_SPLIT=0 $CB --chain $FULL --clips 4096 --synth mixed | line fused
JB_LANE_SPLIT=1 $CB --chain $FULL --clips 16384 --synth mixed | line split
JB_LANE_SPLIT=0 $CB --chain $FULL --clips 16384 --synth mixed | line fused
