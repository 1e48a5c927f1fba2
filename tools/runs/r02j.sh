#!/bin/bash
cd /root/repo
V=juicy-audio-plugins_b200/build/variants
python -m pytest tests/test_exact_math.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02j_pytest.log
python -m pytest tests/test_gpu_population.py -m gpu -x -q -k "c2" > gpurun_out/r02j_pop.log 2>&1; echo "pop rc=$?"; tail -2 gpurun_out/r02j_pop.log
CB="python tools/chain_bench.py --steps 5 --warmup 2 --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum"
fmt() { grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-40s %8.3f ms [%s]' % (sys.argv[1], d['ms_per_render'], d['path']))
" "$1"; }
{
for lib in default noarrive rollch rollch20; do
  if [ $lib = default ]; then L=; else L="JUICY_BATCH_LIB=$V/libjb_$lib.so"; fi
  for iso in 0 1; do
    env $L JB_CO_ISOLATE=$iso $CB 2>&1 | fmt "$lib exact iso=$iso"
    env $L JB_CO_ISOLATE=$iso $CB --math fast 2>&1 | fmt "$lib fast iso=$iso"
  done
done
python tools/chain_bench.py --steps 5 --warmup 2 --chain JuicyPunch,JuicyWidth --clips 4096 --synth mixed 2>&1 | fmt "default exact mixed clips"
python tools/chain_bench.py --steps 5 --warmup 2 --chain JuicyWidth --clips 4096 --synth drum 2>&1 | fmt "Width alone"
python tools/chain_bench.py --steps 5 --warmup 2 --chain JuicyInfer --clips 4096 --synth drum 2>&1 | fmt "Infer alone"
python tools/chain_bench.py --steps 5 --warmup 2 --chain JuicyPunch --clips 4096 --synth drum --math exact 2>&1 | fmt "Punch alone exact"
} | tee gpurun_out/r02j_bench.txt
