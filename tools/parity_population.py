#!/usr/bin/env python
"""Population-level parity report (VERDICT r01 task 1): every clip of C2, C3 (each material and clip mod 5), C4 and a
32768-clip C5 shard (default material and a resonant one) against the CPU oracle; writes one JSON document.

  python tools/parity_population.py [--out gpurun_out/parity_population.json] [--only C2,C4] [--math auto]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import population  # noqa: E402
from conftest import load_juicy_batch  # noqa: E402

FULL_CHAIN = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]
MATERIALS = ("gel", "metal", "wood", "plastic", "flesh")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_population.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--math", default="auto")
    ap.add_argument("--limit", type=int, default=0, help="check only the first N clips of each configuration (all are rendered)")
    args = ap.parse_args()
    jb = load_juicy_batch()
    limit = args.limit or None
    runs = [("C2", dict(chain=["JuicyPunch", "JuicyWidth"], n_clips=4096, synth="drum")),
            ("C2-mixed", dict(chain=["JuicyPunch", "JuicyWidth"], n_clips=4096, synth="mixed")),
            ("C3-mod5", dict(chain=["JuicyTexture"], n_clips=8192, synth="impulse", per_clip={"slot": 0, "id": "material", "mod": 5}))]
    for m, name in enumerate(MATERIALS):
        runs.append(("C3-" + name, dict(chain=["JuicyTexture"], n_clips=8192, synth="impulse", params={0: {"material": float(m)}})))
    runs += [("C4", dict(chain=["JuicyInfer"], n_clips=65536, synth="mixed")),
             ("C5", dict(chain=FULL_CHAIN, n_clips=32768, synth="mixed")),
             ("C5-wood", dict(chain=FULL_CHAIN, n_clips=32768, synth="mixed", params={2: {"material": 2.0}})),
             ("SatWidth-mixed", dict(chain=["JuicySaturator", "JuicyWidth"], n_clips=8192, synth="mixed"))]
    only = [s for s in args.only.split(",") if s]
    out = {"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "math": args.math, "results": []}
    for name, kw in runs:
        if only and name not in only:
            continue
        t0 = time.time()
        try:
            r = population.run_population(jb, math=args.math, limit_clips=limit, name=name, **kw)
        except Exception as exc:
            r = {"config": name, "error": repr(exc)}
        r["wall_seconds"] = time.time() - t0
        out["results"].append(r)
        print(json.dumps(r), flush=True)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
    bad = [r["config"] for r in out["results"] if r.get("error") or r.get("clips_over_sample_tol") or r.get("clips_over_record_tol")]
    print("configurations with violations:", bad or "none")
    return 0


if __name__ == "__main__":
    sys.exit(main())
