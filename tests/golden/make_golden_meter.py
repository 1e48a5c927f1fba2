"""Generates tests/golden/golden_meter_v1.npz from the reference's OWN JuicyMeterPanel
(src/shared/JuicyMeterPanel.cpp compiled unmodified into oracle/_ref/libjuicy_ref_MeterPanel.so by
oracle/Makefile) -- run in the build container, where /root/reference exists:

    python tests/golden/make_golden_meter.py

Inputs: the per-block record histories of every golden case (tests/golden/golden_v1.npz, themselves
made by the reference's processBlock), plus seeded synthetic record sequences that leave [0, 1]
(exercising updateStats' clamp) and have zero pre/post scores (exercising setMetrics' fallback to
`score`).  Stored: the input sequences that are not already in golden_v1.npz, and the 40 numbers
the panel holds afterwards (layout: oracle/ref_meter_harness.cpp)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refhost  # noqa: E402
from cases import GOLDEN_CASES, load_golden  # noqa: E402


def synthetic_sequences():
    rng = np.random.default_rng(0x4A554943)
    seqs = {}
    for n in (1, 2, 7, 94, 938):
        r = rng.uniform(-0.2, 1.2, (n, 16)).astype(np.float32)
        r[:, 0:3] = rng.uniform(0.0, 100.0, (n, 3)).astype(np.float32)
        r[::3, 1] = 0.0
        r[1::4, 2] = 0.0
        seqs["synthetic/%d" % n] = r
    seqs["synthetic/empty"] = np.zeros((0, 16), dtype=np.float32)
    return seqs


def main():
    assert refhost.meter_available(), "build oracle/_ref first: make -C oracle ref"
    z, _ = load_golden()
    store = {}
    count = 0
    for case in GOLDEN_CASES:
        for slot in range(len(case["chain"])):
            key = "hist/%s/%d" % (case["name"], slot)
            store["meter/" + key] = refhost.meter_run(z[key])
            count += 1
    for key, seq in synthetic_sequences().items():
        store["in/" + key] = seq
        store["meter/" + key] = refhost.meter_run(seq)
        count += 1
    path = os.path.join(HERE, "golden_meter_v1.npz")
    np.savez_compressed(path, **store)
    print("wrote %s: %d sequences, %.1f KB" % (path, count, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
