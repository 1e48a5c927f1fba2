#!/bin/bash
# session 4, call 16: JuicyCohere on the few-streams kernel -- parity tests, then timing against the lane kernels
cd /root/repo
python -m pytest tests/test_gpu_solo.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 5
CB="python tools/chain_bench.py --steps 3 --warmup 1 --synth mixed"
for c in 1 16 148 592 888; do
  for path in auto lane; do
    echo -n "Cohere $c clips path=$path: "; $CB --chain JuicyCohere --clips $c --path $path | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"
  done
done | tee gpurun_out/r03i_cohere_solo.txt
for c in 148 592; do echo -n "Saturator $c clips (solo): "; $CB --chain JuicySaturator --clips $c | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f ms' % d['ms_per_render'])"; done | tee -a gpurun_out/r03i_cohere_solo.txt
