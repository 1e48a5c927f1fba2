#!/bin/bash
# session 4, call 3: jb_process_host geometry, interleaved rounds (box drift); C5 shard with two / three plugins side by side
cd /root/repo
X='JB_HOST_MIN_SLICE_BLOCKS=1,JB_HOST_TAPER=0|JB_HOST_MIN_SLICE_BLOCKS=2,JB_HOST_TAPER=0|JB_HOST_MIN_SLICE_BLOCKS=3,JB_HOST_TAPER=0|JB_HOST_MIN_SLICE_BLOCKS=3,JB_HOST_TAPER=1|JB_HOST_MIN_SLICE_BLOCKS=4,JB_HOST_TAPER=1|JB_HOST_MIN_SLICE_BLOCKS=6,JB_HOST_TAPER=1'
python tools/e2e_sweep.py --chain full --clips 32768 --floor --reps 1 --rounds 4 --pass-mib 32768,8192 --slice-mib 96 --extra "$X" > gpurun_out/r02v_c5.txt 2> gpurun_out/r02v_c5.err; echo "c5 rc=$?"
X2='JB_HOST_TAPER=0|JB_HOST_TAPER=1'
python tools/e2e_sweep.py --chain JuicyPunch,JuicyWidth --synth drum --clips 4096 --floor --reps 2 --rounds 5 --pass-mib 32768 --slice-mib 64,96,128 --extra "$X2" > gpurun_out/r02v_c2.txt 2> gpurun_out/r02v_c2.err; echo "c2 rc=$?"
cat gpurun_out/r02v_c5.txt gpurun_out/r02v_c2.txt | cut -c1-420; tail -n 3 gpurun_out/r02v_c5.err gpurun_out/r02v_c2.err
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
CB="python tools/chain_bench.py --steps 3 --warmup 1 --chain $FULL --clips 32768 --synth mixed --inplace"
{
echo "default"; $CB
for k in 2 3 4 8; do echo "pipeline K=$k"; JB_CHAIN_PIPELINE=1 JB_PIPE_SEGMENTS=$k $CB; done
echo "texture pair"; JB_PAIR_LIMIT_TEXTURE=65536 $CB
echo "fast default"; $CB --math fast
echo "fast pipeline K=2"; JB_CHAIN_PIPELINE=1 JB_PIPE_SEGMENTS=2 $CB --math fast
} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   %.2f ms' % d['ms_per_render'])
    else: print(l)
" | tee gpurun_out/r02v_pipe.txt
