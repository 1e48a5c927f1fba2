python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_kernel" 2>&1 | tail -15
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-12s %6d %-16s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
for pm in 0 1; do
  export JB_PAIR=$pm
  for m in 0 1 2 3; do
    $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=$m | line "pair=$pm m$m"
  done
  for c in 16384 32768 65536; do
    $CB --chain JuicyTexture --clips $c --synth impulse --param 0:material=0 | line "pair=$pm gel"
  done
  for c in 8192 16384 32768 65536; do
    $CB --chain JuicySaturator --clips $c --synth sweep --math fast | line "pair=$pm fast"
    $CB --chain JuicySaturator --clips $c --synth sweep --math exact | line "pair=$pm exact"
    $CB --chain JuicyPunch --clips $c --synth drum --math fast | line "pair=$pm fast"
    $CB --chain JuicyPunch --clips $c --synth drum --math exact | line "pair=$pm exact"
  done
done
