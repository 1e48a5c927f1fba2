#!/usr/bin/env python
"""chain_bench.py -- device-resident timing of any plugin chain / clip count / kernel path through the C ABI.

Not the driver's bench (bench.py is): this is the survey tool behind the per-config tables in DESIGN.md §6 and
profiles/.  No torch: device memory, synthetic clips and CUDA-event kernel timing all come from libjuicy_batch.so.

  python tools/chain_bench.py --chain JuicyInfer --clips 65536 --synth mixed --path lane
prints one JSON line: launch ms (mean of --steps renders after --warmup), channel-samples/s, algorithmic GB/s
(8 B per channel-sample, 4 B for a read-only chain, + 64 B per clip-block-plugin) and its
fraction of the measured HBM peak.
"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "juicy-audio-plugins_b200")


def load_juicy_batch():
    spec = importlib.util.spec_from_file_location("juicy_batch", os.path.join(PKG, "juicy_batch.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["juicy_batch"] = mod
    spec.loader.exec_module(mod)
    return mod


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain", default="JuicyPunch,JuicyWidth")
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=48000)
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--synth", default="drum")
    ap.add_argument("--path", default="auto", choices=("auto", "lane", "coop"))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--param", action="append", default=[], help="slot:id=value")
    ap.add_argument("--inplace", action="store_true")
    ap.add_argument("--clipranges", action="store_true")
    ap.add_argument("--math", default="auto", choices=("auto", "exact", "fast"))
    ap.add_argument("--clipmod", default=None, help="id=K: clip c gets parameter id = c mod K in slot 0 (per-clip parameter sets)")
    args = ap.parse_args()

    jb = load_juicy_batch()
    chain = args.chain.split(",")
    n_clips, n = args.clips, args.samples
    count = n_clips * 2 * n
    d_in = jb.DeviceBuffer(count * 4)
    d_out = d_in if args.inplace else jb.DeviceBuffer(count * 4)
    jb.synth_fill_device(d_in.ptr.value, args.synth, 0, n_clips, 2, n, 48000.0, device=0, stream=0)
    eng = jb.BatchProcessor(chain, n_clips, device=0)
    for p in args.param:
        slot, rest = p.split(":", 1)
        pid, val = rest.split("=")
        eng.setParameter(pid, float(val), int(slot))
    if args.clipmod:
        pid, k = args.clipmod.split("=")
        k = int(k)
        if args.clipranges:   # K contiguous ranges instead of c mod K
            per = (n_clips + k - 1) // k
            for j in range(k):
                eng.setParameterClips(pid, float(j), j * per, min(per, n_clips - j * per), 0)
        else:
            for c in range(n_clips):
                eng.setParameterClips(pid, float(c % k), c, 1, 0)
    eng.prepareToPlay(48000.0, args.block)
    eng.set_path(args.path)
    eng.set_math_mode(args.math)
    for _ in range(args.warmup):
        eng.reset()
        eng.process_device(d_in.ptr.value, d_out.ptr.value, n)
    eng.synchronize()
    eng.kernel_time_ms()
    for _ in range(args.steps):
        eng.reset()
        eng.process_device(d_in.ptr.value, d_out.ptr.value, n)
    ms, launches = eng.kernel_time_ms()
    coop, lane = eng.path_launches()
    ms_per = ms / max(launches, 1) * (launches / args.steps)   # a render may be several launches
    read_only = chain == ["JuicyInfer"] and abs(eng.getRawParameterValue("trim") if "trim" in
                                                  [q["id"] for q in eng.parameterInfo(0)] else 0.0) < 1e-9
    n_blocks = (n + args.block - 1) // args.block
    alg = (4.0 if read_only else 8.0) * count + 64.0 * n_clips * n_blocks * len(chain)
    gbs = alg / (ms_per / 1e3) / 1e9
    rec = eng.getLatestMetrics(len(chain) - 1)
    print(json.dumps({"chain": chain, "clips": n_clips, "samples": n, "block": args.block, "synth": args.synth,
                      "path": "coop" if coop else "lane", "ms_per_render": ms_per, "launches_per_render": launches / args.steps,
                      "ch_samples_per_s": count / (ms_per / 1e3), "alg_GBs": gbs, "frac_of_measured_hbm": gbs / peak_gbs(),
                      "alg_bytes": alg, "mean_juiciness": float(rec[:, 13].mean())}))
    eng.close()


if __name__ == "__main__":
    sys.exit(main())
