CB="python tools/chain_bench.py --steps 3 --warmup 1 --path lane"
for v in d8 d10 d14 d24; do
  if [ $v != base ]; then export JUICY_BATCH_LIB=$PWD/juicy-audio-plugins_b200/variants/libjb_$v.so; fi
  for args in "--chain JuicySaturator --clips 8192 --synth sweep" "--chain JuicyTexture --clips 8192 --synth impulse --param 0:material=0"; do
    for ip in "" "--inplace"; do
      $CB $args $ip | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v %-70s %-10s %8.2f ms' % ('$args', '$ip', d['ms_per_render']))"
    done
  done
done
