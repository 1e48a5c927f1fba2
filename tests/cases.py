"""Shared case tables for the parity tests and the golden-vector generator."""
import numpy as np

from conftest import load_juicy_batch

N_SAMPLES = 4400        # 8 x 512 + 304: ragged last block
SAMPLE_RATE = 48000.0
BLOCK = 512
SEED = 0x4A554943

PLUGINS = ("JuicyInfer", "JuicyPunch", "JuicySaturator", "JuicyWidth", "JuicyCohere", "JuicyTexture", "JuicyMotion")
MATERIALS = ("gel", "metal", "wood", "plastic", "flesh")
FULL_CHAIN = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]

# clip ids: drum clips are chosen so the hit starts early inside the 4400-sample window
CLIP_FOR = {"sweep": 0, "noise": 1, "impulse": 2, "drum": 9}


def _case(name, chain, inp, clip=None, programs=None, params=None):
    return {"name": name, "chain": list(chain), "input": inp, "clip": CLIP_FOR[inp] if clip is None else clip,
            "programs": programs, "params": params}


def _build():
    cases = []
    for p in PLUGINS:
        if p == "JuicyTexture":
            continue
        for inp in ("sweep", "noise", "impulse", "drum"):
            cases.append(_case("%s-%s" % (p, inp), [p], inp))
    for m, mat in enumerate(MATERIALS):
        for inp in ("impulse", "drum"):
            cases.append(_case("JuicyTexture-%s-%s" % (mat, inp), ["JuicyTexture"], inp, params={0: {"material": float(m)}}))
    for p, progs in (("JuicySaturator", (1, 3)), ("JuicyPunch", (1, 4)), ("JuicyWidth", (1, 4)), ("JuicyInfer", (2, 4))):
        for g in progs:
            cases.append(_case("%s-program%d-drum" % (p, g), [p], "drum", programs={0: g}))
    cases.append(_case("JuicyCohere-learn-noise", ["JuicyCohere"], "noise", params={0: {"learn": 1.0, "match": 0.9}}))
    cases.append(_case("JuicyMotion-deep-drum", ["JuicyMotion"], "drum",
                       params={0: {"microvar": 0.9, "motiondepth": 1.7, "budget": 0.9}}))
    cases.append(_case("JuicyTexture-wood-tuned-noise", ["JuicyTexture"], "noise",
                       params={0: {"material": 2.0, "tailshape": 0.8, "damping": 0.2, "weight": 0.7, "texture": 0.3}}))
    cases.append(_case("chain-punch-width-drum", ["JuicyPunch", "JuicyWidth"], "drum"))
    cases.append(_case("chain-full-drum", FULL_CHAIN, "drum"))
    cases.append(_case("chain-full-noise", FULL_CHAIN, "noise"))
    return cases


GOLDEN_CASES = _build()


def case_input(case, n_samples=N_SAMPLES):
    """[2][n] float32 input of a case, from the library's host generator (needs no GPU)."""
    jb = load_juicy_batch()
    return jb.synth_clips(case["input"], case["clip"], 1, n_samples, 2, SAMPLE_RATE, SEED)[0]


def apply_case_settings(engine, case):
    """Programs first, then individual parameters -- the order refhost.run_chain uses."""
    for slot in range(len(case["chain"])):
        if case.get("programs") and slot in case["programs"]:
            engine.setCurrentProgram(case["programs"][slot], slot)
        if case.get("params") and slot in case["params"]:
            for k, v in case["params"][slot].items():
                engine.setParameter(k, v, slot)


def load_golden():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta
