#!/bin/bash
# session 3, run 1: exact tanh specialisations + voted punch_chunk
cd /root/repo
python -m pytest tests/test_exact_math.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02i_pytest.log
python -m pytest tests/test_gpu_population.py -m gpu -x -q -k "c2 or c5_shard_32768" > gpurun_out/r02i_pop.log 2>&1; echo "pop rc=$?"; tail -3 gpurun_out/r02i_pop.log
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
CB="python tools/chain_bench.py --steps 5 --warmup 2"
{
$CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum
$CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum --math fast
$CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth mixed
$CB --chain $FULL --clips 32768 --synth mixed --inplace
JB_PAIR=1 $CB --chain $FULL --clips 32768 --synth mixed --inplace
$CB --chain JuicyTexture --clips 32768 --synth mixed --inplace
JB_PAIR=1 $CB --chain JuicyTexture --clips 32768 --synth mixed --inplace
$CB --chain JuicySaturator --clips 65536 --synth mixed --inplace --math exact
$CB --chain JuicyPunch --clips 65536 --synth mixed --inplace --math exact
} 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-50s %6d %8.2f ms %5.1f%% [%s]' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], d['ms_per_render'], 100*d['frac_of_measured_hbm'], d['path']))
" | tee gpurun_out/r02i_bench.txt
