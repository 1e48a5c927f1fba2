"""Generates tests/golden/golden_v1.npz from the reference's OWN C++ (oracle/_ref, i.e.
/root/reference/src/... compiled unmodified by oracle/Makefile) -- run in the build
container, where /root/reference exists:

    python tests/golden/make_golden.py

Every case stores the output samples and the per-block metric records of one clip; the
inputs are stored once per (kind, clip).  tests/test_oracle_port.py pins the C
restatement (oracle/juicy_oracle.c) to these bit for bit, and tests/test_gpu_parity.py
checks the CUDA engine against them within the stated tolerances.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refhost  # noqa: E402
from cases import GOLDEN_CASES, N_SAMPLES, SAMPLE_RATE, BLOCK, case_input  # noqa: E402


def main():
    assert refhost.available(), "build oracle/_ref first: make -C oracle ref"
    store = {}
    meta = {"n_samples": N_SAMPLES, "sample_rate": SAMPLE_RATE, "block": BLOCK, "cases": []}
    for case in GOLDEN_CASES:
        x = case_input(case)
        key_in = "in/%s/%d" % (case["input"], case["clip"])
        store[key_in] = x
        out, hists = refhost.run_chain(case["chain"], x, sample_rate=SAMPLE_RATE, block_size=BLOCK,
                                       programs=case.get("programs"), params=case.get("params"))
        store["out/" + case["name"]] = out
        for slot, h in enumerate(hists):
            store["hist/%s/%d" % (case["name"], slot)] = h
        meta["cases"].append(case["name"])
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **store)
    print("wrote %s: %d cases, %.1f KB" % (path, len(GOLDEN_CASES), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
