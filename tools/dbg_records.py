#!/usr/bin/env python
"""Debug aid: render a configuration on the GPU (twice, to see run-to-run differences), then print -- for the given clips --
every (block, plugin) record that differs from the CPU oracle by more than 0.01, field by field.
  python tools/dbg_records.py --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=0 --look 8038,8057"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_juicy_batch  # noqa: E402
from oracle import port  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain", default="JuicyTexture")
    ap.add_argument("--clips", type=int, default=8192)
    ap.add_argument("--samples", type=int, default=48000)
    ap.add_argument("--synth", default="impulse")
    ap.add_argument("--param", action="append", default=[])
    ap.add_argument("--look", default="0")
    ap.add_argument("--math", default="auto")
    args = ap.parse_args()
    jb = load_juicy_batch()
    chain = args.chain.split(",")
    params = {}
    for p in args.param:
        slot, kv = p.split(":")
        k, v = kv.split("=")
        params.setdefault(int(slot), {})[k] = float(v)
    n, nc = args.samples, args.clips
    nb = (n + 511) // 512
    clip_bytes = 2 * n * 4
    d_in = jb.DeviceBuffer(nc * clip_bytes)
    d_out = jb.DeviceBuffer(nc * clip_bytes)
    jb.synth_fill_device(d_in.ptr.value, args.synth, 0, nc, 2, n)
    hists = []
    for run in range(2):
        eng = jb.BatchProcessor(chain, nc)
        for slot, kv in params.items():
            for k, v in kv.items():
                eng.setParameter(k, v, slot)
        eng.set_math_mode(args.math)
        eng.enableHistory(nb)
        eng.prepareToPlay(48000.0, 512)
        eng.process_device(d_in.ptr.value, d_out.ptr.value, n)
        eng.synchronize()
        hists.append([eng.getHistory(s, 0, nb) for s in range(len(chain))])
        eng.close()
    for s in range(len(chain)):
        same = np.array_equal(hists[0][s], hists[1][s])
        print("slot %d: run 0 == run 1: %s" % (s, same))
        if not same:
            d = np.abs(hists[0][s] - hists[1][s]).max(axis=(0, 2))
            print("   clips differing between runs:", np.nonzero(d > 0)[0][:40])
    for c in [int(x) for x in args.look.split(",")]:
        x = np.empty((2, n), dtype=np.float32)
        jb._check(jb.lib().jb_copy_to_host(0, x.ctypes.data, d_in.ptr.value + c * clip_bytes, clip_bytes))
        g = np.empty((2, n), dtype=np.float32)
        jb._check(jb.lib().jb_copy_to_host(0, g.ctypes.data, d_out.ptr.value + c * clip_bytes, clip_bytes))
        ref, rh = port.run_chain(chain, x, params=params or None)
        print("clip %d: max |gpu-ref|/peak = %.3e" % (c, np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-30)))
        for s in range(len(chain)):
            for b in range(nb):
                d = np.abs(hists[0][s][b, c] - rh[s][b])
                if d.max() > 0.01:
                    blk = ref[:, b * 512:(b + 1) * 512]
                    print("  slot %d block %d: |ref out| max %.3e rms %.3e" % (s, b, np.abs(blk).max(), np.sqrt((blk.astype(np.float64) ** 2).mean())))
                    for f, name in enumerate(jb.RECORD_FIELDS):
                        if d[f] > 1e-4:
                            print("      %-18s gpu %.6f  ref %.6f" % (name, hists[0][s][b, c, f], rh[s][b, f]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
