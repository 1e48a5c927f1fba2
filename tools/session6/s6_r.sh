python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined" 2>&1 | tail -5
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane --inplace"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-44s %6d %-12s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
for pm in 0 1; do
  export JB_CHAIN_PIPELINE=$pm
  for c in 1024 4096 8192 16384 32768; do
    $CB --chain $FULL --clips $c --synth mixed | line "pipe=$pm"
  done
  $CB --chain JuicySaturator,JuicyWidth,JuicyCohere,JuicyInfer --clips 16384 --synth mixed | line "pipe=$pm"
  $CB --chain JuicySaturator,JuicyCohere --clips 65536 --synth mixed | line "pipe=$pm"
  $CB --chain $FULL --clips 8192 --synth mixed --param 2:material=2 | line "pipe=$pm wood"
done
