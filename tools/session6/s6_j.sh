CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-12s %6d %-10s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for v in sel sel64; do
  export JUICY_BATCH_LIB=$PWD/juicy-audio-plugins_b200/variants/libjb_$v.so
  for p in JuicyInfer JuicySaturator JuicyCohere JuicyWidth JuicyPunch; do
    $CB --chain $p --clips 65536 --synth mixed | line $v
  done
  $CB --chain JuicyInfer --clips 32768 --synth mixed | line $v
  $CB --chain JuicySaturator --clips 32768 --synth mixed | line $v
done
