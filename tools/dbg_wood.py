"""dbg_wood.py -- how far a chain drifts from the oracle when Punch / Saturator run their MUFU-based tanh / pow in front of
each Texture material (the measurement behind csrc/jb_libm.h; output kept in profiles/r01_s6_chain_sensitivity.txt).
Force the fast routines with eng.set_math_mode("fast") to reproduce it now that auto mode switches to the exact ones."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from conftest import load_juicy_batch
from oracle import port
jb = load_juicy_batch()
FULL = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]
def run(chain, settings, n_clips=12, n=2*512+128, kind="mixed"):
    clips = jb.synth_clips(kind, 11, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    for slot, pid, v in settings:
        if pid == "__program__": eng.setCurrentProgram(int(v), slot)
        else: eng.setParameter(pid, v, slot)
    eng.prepareToPlay(48000.0, 512)
    out = eng.processBlock(clips)
    eng.close()
    worst = []
    for c in range(n_clips):
        plugs = [port.PortPlugin(p, 2, 48000.0, 512) for p in chain]
        for slot, pid, v in settings:
            if pid == "__program__": plugs[slot].set_program(int(v))
            else: plugs[slot].set_param(pid, v)
        cur = clips[c]
        for p in plugs:
            p.prepare(); cur, _ = p.process(cur)
        worst.append(float(np.abs(out[c]-cur).max()/max(np.abs(cur).max(),1e-30)))
    print("%-60s %s max %.2e" % ("+".join(x[5:] for x in chain), settings, max(worst)), ["%.1e" % w for w in worst])
for m in range(5):
    run(FULL, [(2, "material", float(m))])
run(FULL, [(0, "__program__", 2), (2, "material", 2.0)])
run(["JuicyPunch", "JuicyTexture"], [(0, "__program__", 2), (1, "material", 2.0)])
run(["JuicyPunch", "JuicyTexture"], [(1, "material", 2.0)])
run(["JuicySaturator", "JuicyTexture"], [(1, "material", 2.0)])
run(["JuicyTexture"], [(0, "material", 2.0)])
run(["JuicyTexture"], [(0, "material", 3.0)])
