python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_streaming" 2>&1 | tail -12
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane --math fast"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-12s %6d %-16s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for t in 0 1; do
  export JB_TILE=$t
  for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer; do
    $CB --chain $p --clips 65536 --synth mixed | line "tile=$t"
  done
  $CB --chain JuicySaturator --clips 32768 --synth mixed | line "tile=$t"
  $CB --chain JuicySaturator --clips 16384 --synth mixed | line "tile=$t"
  $CB --chain JuicyInfer --clips 32768 --synth mixed | line "tile=$t"
  $CB --chain JuicyInfer --clips 16384 --synth mixed | line "tile=$t"
  $CB --chain JuicyInfer --clips 65536 --synth mixed --inplace | line "tile=$t inplace"
done
