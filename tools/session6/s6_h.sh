CB="python tools/chain_bench.py --steps 3 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-40s %6d %-14s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
$CB --chain JuicyInfer --clips 65536 --synth mixed | line "single"
$CB --chain JuicyInfer --clips 65536 --synth mixed | line "single again"
JB_LANE_GENERIC=1 $CB --chain JuicyInfer --clips 65536 --synth mixed | line "generic"
$CB --chain JuicyInfer --clips 65536 --synth mixed --inplace | line "single inplace"
$CB --chain JuicyInfer --clips 65536 --synth noise | line "single noise"
$CB --chain JuicyInfer --clips 65536 --synth sweep | line "single sweep"
$CB --chain JuicyInfer --clips 32768 --synth mixed | line "single 32768"
$CB --chain JuicySaturator --clips 65536 --synth sweep | line "sat sweep"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
