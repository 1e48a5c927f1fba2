#!/bin/bash
# session 4, call 11 (2 GPUs): the bench line at N = 2 as the driver launches it, engine arm and reference arm; the cross-device sharding test
cd /root/repo
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r03d_bench_n2.json 2> gpurun_out/r03d_bench_n2.err; echo "bench n2 rc=$?"; tail -n 4 gpurun_out/r03d_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r03d_bench_n2.json').read().strip().splitlines()[-1])
r=d['roofline']; e=d['e2e']
print('n', d['n_gpus'], d['config']['name'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], r['kernel'][:30], 'step', r['step']['frac'], 'e2e med/mean/min', e['ms_per_step'], e['ms_per_step_mean'], e['ms_per_step_min'], 'floor', e['pcie_floor_ms'], 'pcm', d['e2e_pcm16']['ms_per_step'], 'launches', d['gpu_launches'])
"
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3) > gpurun_out/r03d_bench_n2_reference.json 2> gpurun_out/r03d_bench_n2_reference.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/r03d_bench_n2_reference.json
python -m pytest tests/test_gpu_sharding.py -m gpu -x -q 2>&1 | tail -n 3
