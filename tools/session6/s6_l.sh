python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane"
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-40s %6d %-16s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm']))
" "$1"; }
$CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 | line "C3 mod5"
for m in 0 1 2 3 4; do $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=$m | line "C3 m$m"; done
$CB --chain JuicyInfer --clips 65536 --synth mixed | line "C4"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
$CB --chain $FULL --clips 32768 --synth mixed | line "C5 shard auto(exact)"
$CB --chain $FULL --clips 32768 --synth mixed --math fast | line "C5 shard fast"
$CB --chain $FULL --clips 4096 --synth mixed | line "auto(exact)"
$CB --chain $FULL --clips 8192 --synth mixed | line "auto(exact)"
$CB --chain JuicySaturator --clips 1 --samples 480000 --synth sweep | line "C1"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s6b.json 2> gpurun_out/bench_s6b.err; tail -c 300 gpurun_out/bench_s6b.json
