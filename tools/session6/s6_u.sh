CB="python tools/chain_bench.py --steps 3 --warmup 1 --path lane"
for args in "--chain JuicyTexture --clips 8192 --synth impulse --param 0:material=0" "--chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5" "--chain JuicySaturator --clips 65536 --synth sweep" "--chain JuicySaturator --clips 8192 --synth sweep" "--chain JuicyMotion --clips 8192 --synth drum" "--chain JuicyWidth --clips 65536 --synth mixed"; do
  for ip in "" "--inplace"; do
    $CB $args $ip | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-75s %-10s %8.2f ms' % ('$args', '$ip', d['ms_per_render']))"
  done
done
