CB="python tools/chain_bench.py --steps 3 --warmup 1 --path lane"
for pm in 0 1; do
for args in "--chain JuicyTexture --clips 8192 --synth impulse --param 0:material=0" "--chain JuicySaturator --clips 8192 --synth sweep" "--chain JuicyPunch --clips 8192 --synth drum" "--chain JuicyCohere --clips 8192 --synth mixed"; do
  for ip in "" "--inplace"; do
    JB_PAIR=$pm $CB $args $ip | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pair=$pm %-75s %-10s %8.2f ms' % ('$args', '$ip', d['ms_per_render']))"
  done
done
done
