#!/bin/bash
# session 4, call 20: which kernels does a JuicyCohere render of 148 clips launch, and how long is each
cd /root/repo
for p in JuicyCohere JuicyInfer; do
echo "== $p"
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,smsp__inst_executed.sum --clock-control none python tools/chain_bench.py --steps 1 --warmup 1 --synth mixed --chain $p --clips 148 2>&1 | grep -E "^\s+(void )?<unnamed>|^\s+jb_|gpu__time|registers|shared_mem|inst_executed" | grep -v fill | cut -c1-150 | tail -n 14
done
