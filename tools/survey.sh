#!/bin/bash
# survey.sh -- device-resident timing tables through tools/chain_bench.py (replaces the round-1 one-off scripts).
#   tools/survey.sh final                 every plugin alone at 65536 / 8192 clips, then C1..C5 (the DESIGN.md §6 table)
#   tools/survey.sh plugins CLIPS...      every plugin alone at the given clip counts
#   tools/survey.sh ab VAR v1 v2 -- ARGS  chain_bench ARGS once per value of environment variable VAR (A/B of a kernel knob,
#                                         e.g.  tools/survey.sh ab JB_LDG256 0 1 -- --chain JuicyCohere --clips 65536 --inplace)
CB="python tools/chain_bench.py --steps 3 --warmup 1"
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
line() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-44s %6d %-26s %8.2f ms %7.1f G ch-samples/s %5.1f%% of HBM  [%s]' % ('+'.join(x.replace('Juicy','') for x in d['chain']), d['clips'], sys.argv[1], d['ms_per_render'], d['ch_samples_per_s']/1e9, 100*d['frac_of_measured_hbm'], d['path']))
" "$1"; }
case "$1" in
final)
  for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer JuicyPunch JuicyMotion JuicyTexture; do
    $CB --chain $p --clips 65536 --synth mixed --inplace | line "65536 in place"
    $CB --chain $p --clips 8192 --synth mixed --inplace | line "8192 in place"
  done
  for p in JuicySaturator JuicyPunch; do $CB --chain $p --clips 65536 --synth mixed --inplace --math exact | line "65536 exact math"; done
  $CB --chain JuicySaturator --clips 1 --samples 480000 --synth sweep | line "C1"
  $CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum | line "C2 (auto = exact)"
  $CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum --math fast | line "C2 fast math"
  $CB --chain JuicyPunch,JuicyWidth --clips 4096 --synth drum --path lane | line "C2 lane kernels"
  $CB --chain JuicyTexture --clips 8192 --synth impulse --clipmod material=5 | line "C3 material = clip mod 5"
  for m in 0 1 2 3 4; do $CB --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=$m | line "C3 material $m"; done
  $CB --chain JuicyInfer --clips 65536 --synth mixed --inplace | line "C4 in place"
  $CB --chain JuicyInfer --clips 65536 --synth mixed | line "C4 out of place"
  $CB --chain $FULL --clips 32768 --synth mixed --inplace | line "C5 shard (auto = exact)"
  $CB --chain $FULL --clips 32768 --synth mixed --inplace --math fast | line "C5 shard fast math"
  $CB --chain $FULL --clips 32768 --synth mixed --inplace --param 2:material=2 | line "C5 shard, Texture wood"
  $CB --chain $FULL --clips 4096 --synth mixed --inplace | line "7-plugin chain, 4096 clips"
  ;;
plugins)
  shift
  for c in "$@"; do for p in JuicySaturator JuicyCohere JuicyWidth JuicyInfer JuicyPunch JuicyMotion JuicyTexture; do
    $CB --chain $p --clips $c --synth mixed --inplace | line "$c in place"; done; done
  ;;
ab)
  var=$2; shift 2; vals=()
  while [ "$1" != "--" ] && [ $# -gt 0 ]; do vals+=("$1"); shift; done; shift
  for v in "${vals[@]}"; do env $var=$v $CB "$@" | line "$var=$v"; done
  ;;
*) sed -n 2,7p "$0";;
esac
