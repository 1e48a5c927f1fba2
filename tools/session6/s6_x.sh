CB="python tools/chain_bench.py --steps 2 --warmup 1 --path lane --inplace --math fast"
for args in "--chain JuicySaturator --synth sweep" "--chain JuicyPunch --synth drum" "--chain JuicyTexture --synth mixed"; do
  for c in 12288 16384 24576 32768; do
    a=$(JB_PAIR=0 $CB $args --clips $c | python -c "import json,sys; print(round(json.loads(sys.stdin.readline())['ms_per_render'],2))")
    b=$(JB_PAIR=1 $CB $args --clips $c | python -c "import json,sys; print(round(json.loads(sys.stdin.readline())['ms_per_render'],2))")
    echo "$args clips $c one-lane $a ms pair $b ms"
  done
done
