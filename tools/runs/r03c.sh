#!/bin/bash
# session 4, call 10: new tests; 16-bit PCM slice sizes on the C5 shard; C4 end to end with the download skipped
cd /root/repo
python -m pytest tests/test_round2_fixes.py tests/test_gpu_pcm16.py tests/test_wav.py -m gpu -x -q 2>&1 | tail -n 4
python tools/e2e_sweep.py --pcm16 --chain full --clips 32768 --reps 1 --rounds 3 --pass-mib 32768 --slice-mib 96 --extra "JB_HOST_MIN_SLICE_BLOCKS=2|JB_HOST_MIN_SLICE_BLOCKS=3|JB_HOST_MIN_SLICE_BLOCKS=4|JB_HOST_MIN_SLICE_BLOCKS=6|JB_HOST_MIN_SLICE_BLOCKS=8" | cut -c1-260
python tools/e2e_sweep.py --chain JuicyInfer --clips 65536 --reps 1 --rounds 3 --pass-mib 32768 --slice-mib 96 | cut -c1-260
