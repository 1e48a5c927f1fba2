python -m pytest tests -x -q -m gpu 2>&1 | tail -6
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-12s %6d %-16s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer JuicyPunch; do
  $CB --chain $p --clips 65536 --synth mixed --math fast | line "65536"
done
$CB --chain JuicySaturator --clips 32768 --synth mixed --math fast | line "32768"
$CB --chain JuicyInfer --clips 32768 --synth mixed | line "32768"
