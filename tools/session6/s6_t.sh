for m in 96 64 48 32 24 16; do
  JB_HOST_SLICE_MIB=$m python bench.py --steps 5 --warmup 3 --no-survey --no-cpu 2>/dev/null > /tmp/b.json
  python -c "
import json
d=json.loads(open('/tmp/b.json').readline()); print('slice MiB $m e2e ms', round(d['e2e']['ms_per_step'],2), 'device ms', round(d['ms_per_step'],3))"
done
python tools/pcie_probe.py 2>&1 | tail -3
