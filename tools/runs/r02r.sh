#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_optin_paths.py tests/test_gpu_fullsize.py tests/test_exact_math.py -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02r_pytest.log
fmt() { grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-52s %8.3f ms %5.1f%% [%s]' % (sys.argv[1], d['ms_per_render'], 100*d['frac_of_measured_hbm'], d['path']))
" "$1"; }
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed --inplace"
{
for c in 32768 65536; do
  $CB --chain JuicySaturator --clips $c --math exact 2>&1 | fmt "Sat exact $c default (pair)"
  for o in 0 1 2; do
    JB_PAIR=0 JB_OCTETS=$o $CB --chain JuicySaturator --clips $c --math exact 2>&1 | fmt "Sat exact $c single octets=$o"
  done
done
$CB --chain JuicyPunch --clips 32768 --math exact 2>&1 | fmt "Punch exact 32768 default (pair)"
JB_PAIR=0 $CB --chain JuicyPunch --clips 32768 --math exact 2>&1 | fmt "Punch exact 32768 single"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
$CB --chain $FULL --clips 32768 2>&1 | fmt "C5 shard exact"
$CB --chain $FULL --clips 32768 --math fast 2>&1 | fmt "C5 shard fast"
for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer JuicyPunch; do
  $CB --chain $p --clips 65536 2>&1 | fmt "$p 65536 (new default)"
done
} | tee gpurun_out/r02r_bench.txt
