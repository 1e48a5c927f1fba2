CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-12s %6d %-22s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for pm in 0 1; do
  export JB_PAIR=$pm
  for syn in impulse mixed drum; do
    for c in 16384 32768; do
      $CB --chain JuicyTexture --clips $c --synth $syn --param 0:material=0 | line "pair=$pm gel $syn"
    done
  done
  $CB --chain JuicyTexture --clips 8192 --synth mixed --param 0:material=0 | line "pair=$pm gel mixed"
  $CB --chain JuicyTexture --clips 8192 --synth mixed --param 0:material=1 | line "pair=$pm metal mixed"
  $CB --chain JuicyTexture --clips 32768 --synth mixed --param 0:material=2 | line "pair=$pm wood mixed"
done
