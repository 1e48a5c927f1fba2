"""Small renders through every kernel variant, for
  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py   (shared-memory hazards: the cooperative kernel's
                                                                       named barriers, mbarriers and cross-warp hand-offs)
  compute-sanitizer --tool synccheck python tools/sanitize_small.py
Logs of the round's runs: profiles/r02_sanitizer_*.log."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from conftest import load_juicy_batch

jb = load_juicy_batch()
FULL = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]


def run(chain, n_clips, n, channels=2, settings=(), per_clip=False, math="auto", block=512):
    clips = jb.synth_clips("mixed", 1, n_clips, n, channels)
    eng = jb.BatchProcessor(chain, n_clips, n_channels=channels)
    for slot, pid, v in settings:
        eng.setParameter(pid, v, slot)
    if per_clip:
        for c in range(n_clips):
            eng.setParameterClips("material", float(c % 5), c, 1, chain.index("JuicyTexture"))
    eng.set_math_mode(math)
    eng.prepareToPlay(48000.0, block)
    eng.enableHistory(16)
    out = eng.processBlock(clips)
    eng.meterStatistics(0)
    eng.close()
    assert np.isfinite(out).all()
    print("ok", "+".join(c[5:] for c in chain), n_clips, n, channels, math, flush=True)


run(FULL, 49, 5 * 512 + 12)                                   # pipelined chain, pair + single kernels, ragged tail
run(FULL, 33, 3 * 512 + 7)                                    # non-vector path (n % 4 != 0)
run(FULL, 64, 4 * 512, settings=[(2, "material", 2.0)])       # exact math, waveguide material
run(["JuicyTexture"], 67, 3 * 512 + 40, per_clip=True)        # clip maps, concurrent launches
run(["JuicySaturator", "JuicyWidth", "JuicyCohere", "JuicyInfer"], 64, 5 * 512 + 12)   # JB_TILE=1 from the environment
run(FULL, 37, 3 * 512 + 50, channels=1, math="fast")          # mono kernel
run(["JuicyPunch", "JuicyWidth"], 96, 4 * 512)                # cooperative kernel, exact-math instantiation (auto)
run(["JuicyPunch", "JuicyWidth", "JuicyInfer"], 70, 3 * 512 + 256, math="fast")   # cooperative kernel, fast math, ragged groups
run(["JuicyWidth"], 40, 2 * 512 + 128, block=256)             # cooperative kernel, Width alone, one step per block
run(["JuicySaturator", "JuicyTexture"], 21, 2 * 512 + 64, channels=1, settings=[(1, "material", 1.0)])  # mono, exact math
