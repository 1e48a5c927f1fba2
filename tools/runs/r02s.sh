#!/bin/bash
cd /root/repo
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02s_pytest.log
(time python bench.py) > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02s_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
print('C2', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'fast', d['fast_math']['ms_per_step'])
for o in d['other_configs']:
    print(o['config']['name'], o['ms_per_step'], o['roofline']['frac'], 'e2e', o['e2e']['ms_per_step'], o.get('fast_math',{}).get('ms_per_step'))
"
