#!/bin/bash
cd /root/repo
python -m pytest tests/test_exact_math.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02l_pytest.log
python -m pytest tests/test_gpu_population.py -m gpu -x -q -k "c2" > gpurun_out/r02l_pop.log 2>&1; echo "pop rc=$?"; tail -2 gpurun_out/r02l_pop.log
fmt() { grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-40s %8.3f ms [%s]' % (sys.argv[1], d['ms_per_render'], d['path']))
" "$1"; }
CB="python tools/chain_bench.py --steps 5 --warmup 2 --clips 4096 --synth drum"
{
for iso in 0 1; do
  JB_CO_ISOLATE=$iso $CB --chain JuicyPunch,JuicyWidth 2>&1 | fmt "C2 exact iso=$iso"
  JB_CO_ISOLATE=$iso $CB --chain JuicyPunch,JuicyWidth --math fast 2>&1 | fmt "C2 fast iso=$iso"
done
$CB --chain JuicyPunch,JuicyWidth --synth mixed 2>&1 | fmt "C2 exact mixed clips"
$CB --chain JuicyWidth 2>&1 | fmt "Width alone"
$CB --chain JuicyInfer 2>&1 | fmt "Infer alone"
$CB --chain JuicyPunch --math exact 2>&1 | fmt "Punch alone exact"
$CB --chain JuicyPunch 2>&1 | fmt "Punch alone fast"
$CB --chain JuicyPunch,JuicyWidth,JuicyInfer 2>&1 | fmt "Punch+Width+Infer exact"
} | tee gpurun_out/r02l_bench.txt
