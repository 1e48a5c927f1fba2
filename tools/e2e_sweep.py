#!/usr/bin/env python
"""e2e_sweep.py -- jb_process_host's time-slice / pass geometry against the wall clock (one GPU, no torch).

For one workload (chain, clips) the host-buffer render is timed once per (JB_HOST_PASS_MIB, JB_HOST_SLICE_MIB) pair -- the
library reads both on every call -- and once with the cheapest chain on the same buffers (JuicyInfer, trim = 0: a 7 ms
render), which is the copy-bound floor of that geometry.  One line per combination.

  python tools/e2e_sweep.py --chain full --clips 32768 --pass-mib 8192,16384 --slice-mib 96,192,384,768
"""
import argparse
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "juicy-audio-plugins_b200")
FULL = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]


def load_juicy_batch():
    spec = importlib.util.spec_from_file_location("juicy_batch", os.path.join(PKG, "juicy_batch.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["juicy_batch"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain", default="full")
    ap.add_argument("--clips", type=int, default=32768)
    ap.add_argument("--samples", type=int, default=48000)
    ap.add_argument("--synth", default="mixed")
    ap.add_argument("--pass-mib", default="8192")
    ap.add_argument("--slice-mib", default="96")
    ap.add_argument("--extra", default="", help="comma list of NAME=VALUE environment settings swept as a third axis, '|' between alternatives")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--pcm16", action="store_true", help="jb_process_host_pcm16 on 16-bit host buffers (in place) instead of jb_process_host")
    ap.add_argument("--floor", action="store_true", help="also time the JuicyInfer-only render of the same buffers")
    ap.add_argument("--smi-ms", type=int, default=0, help="poll nvidia-smi (clocks, power, throttle reasons) every N ms meanwhile, as bench.py does")
    ap.add_argument("--bind", action="store_true", help="bind to the CPUs NVML reports as local to GPU 0 before allocating, as bench.py does")
    args = ap.parse_args()
    if args.bind:
        import pynvml
        pynvml.nvmlInit()
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0), (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & os.sched_getaffinity(0)
        print(json.dumps({"bind": sorted(cpus)[:4] + ["..."] + sorted(cpus)[-2:], "n": len(cpus), "allowed": len(os.sched_getaffinity(0))}))
        if cpus:
            os.sched_setaffinity(0, cpus)
    smi = None
    if args.smi_ms:
        import subprocess
        smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                                "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                "--format=csv,noheader,nounits", "-lms", str(args.smi_ms)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    jb = load_juicy_batch()
    chain = FULL if args.chain == "full" else args.chain.split(",")
    n_clips, n = args.clips, args.samples
    count = n_clips * 2 * n
    d = jb.DeviceBuffer(count * 4)
    jb.synth_fill_device(d.ptr.value, args.synth, 0, n_clips, 2, n, 48000.0, device=0, stream=0)
    h = jb.PinnedBuffer((n_clips, 2, n))
    h2 = jb.PinnedBuffer((n_clips, 2, n)) if count * 4 <= 4e9 else h
    jb._check(jb.lib().jb_copy_to_host(0, h.array.ctypes.data, d.ptr.value, count * 4))
    d.free()
    engines = [("chain", jb.BatchProcessor(chain, n_clips, device=0))]
    if args.floor:
        engines.append(("floor(Infer)", jb.BatchProcessor(["JuicyInfer"], n_clips, device=0)))
    for _, e in engines:
        e.prepareToPlay(48000.0, 512)
    if args.pcm16:
        import numpy as np
        pcm = h.array.view(np.int16).reshape(-1)[:count]          # the first half of the float buffer, as int16 samples
        step = 1 << 24
        for i in range(0, count, step):                            # any 16-bit content will do for timing: reinterpret the floats' bits
            pass
        def call(e):
            e.process_host_pcm16_ptr(pcm.ctypes.data, pcm.ctypes.data, n)
    else:
        def call(e):
            e.process_host_ptr(h.array.ctypes.data, h2.array.ctypes.data, n)
    extras = [x for x in args.extra.split("|")] if args.extra else [""]
    combos = [(x, p, sl) for x in extras for p in args.pass_mib.split(",") for sl in args.slice_mib.split(",")]
    samples = {}
    keys = ("JB_HOST_PASS_MIB", "JB_HOST_SLICE_MIB", "JB_HOST_TAPER", "JB_HOST_MIN_SLICE_BLOCKS")
    t_start = time.perf_counter()
    # the combinations are walked --rounds times, interleaved, so that drift of the box (other tenants on the host's PCIe /
    # memory path, warm-up) hits all of them alike; every call's time is kept
    for rnd in range(args.rounds):
        for combo in combos:
            extra, p, sl = combo
            for k in keys:
                os.environ.pop(k, None)
            for kv in extra.split(","):
                if kv:
                    k, v = kv.split("=")
                    os.environ[k] = v
            os.environ["JB_HOST_PASS_MIB"] = p
            os.environ["JB_HOST_SLICE_MIB"] = sl
            for name, e in engines:
                if rnd == 0:
                    e.reset()
                    call(e)   # staging allocation of this geometry
                for _ in range(args.reps):
                    e.reset()
                    t0 = time.perf_counter()
                    call(e)
                    samples.setdefault((combo, name), []).append((time.perf_counter() - t0) * 1e3)
    for combo in combos:
        extra, p, sl = combo
        row = {"clips": n_clips, "pass_mib": int(p), "slice_mib": int(sl), "env": extra}
        for name, _ in engines:
            v = sorted(samples[(combo, name)])
            row[name] = {"min_ms": round(v[0], 2), "median_ms": round(v[len(v) // 2], 2), "max_ms": round(v[-1], 2),
                         "GBs_per_dir_at_median": round(count * 4 / v[len(v) // 2] / 1e6, 2),
                         "all_ms": [round(x, 1) for x in samples[(combo, name)]]}
        print(json.dumps(row), flush=True)
    print(json.dumps({"wall_s": round(time.perf_counter() - t_start, 1)}))
    for _, e in engines:
        e.close()
    if smi is not None:
        smi.terminate()
    return 0


if __name__ == "__main__":
    sys.exit(main())
