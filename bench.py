#!/usr/bin/env python
"""bench.py -- channel-samples/s of the JuicySuite hot path on B200 (BASELINE.json metric).

Workloads (BASELINE.json `configs`, shapes of SURVEY.md §8(d)); `--config` picks one.  The default at EVERY N is
  C5 = configs[4], the full 7-plugin chain, 32768 stereo clips per GPU (262144 at 8 GPUs), weak scaling -- the one
  configuration BASELINE.json's metric ("... at 1/2/4/8 B200") spans, so the N = 1, 2, 4, 8 lines are the same workload
  per GPU and their ratio is the scaling.  At N = 1 the other four configurations follow in `other_configs` -- C2 =
  configs[1] (JuicyPunch -> JuicyWidth on 4096 drum-hit clips, the round-1 headline) first, then C1 / C3 / C4 -- each
  with its own device-resident time, roofline, e2e, cpu_baseline and clocks.
Clips are independent plugin-instance chains, so rank r renders its own clips and the data path has no collective;
the per-clip Juiciness records are gathered with ONE ncclAllGather per step issued by the library itself
(jb_gather_records, NCCL resolved with dlopen), inside the timed region.

A step = prepareToPlay-reset + one render of the whole batch, out of place (every input is far larger than the
126 MB L2, so no step sees a warm cache).
  value    : device-resident throughput (inputs already in HBM), CUDA events on the engine's stream, max over ranks
  e2e      : the same render through jb_process_host with pinned HOST buffers, H2D + D2H inside the timed region; every
             step is timed on its own and the figure is the MEDIAN step (max over ranks per step) -- the boxes' host <->
             device path is shared with other tenants and single steps scatter by tens of percent; mean / min / max beside it
  roofline : the dominant kernel's algorithmic bytes per launch / its mean device duration (CUDA events recorded by the
             library on its own stream: around the render, and -- for a chain that renders as one launch per plugin --
             between the plugins' launches, jb_slot_time_ms) vs MEASURED_PEAKS.json; `step` beside it is the whole
             render (all of the chain's launches) against the chain's algorithmic bytes
  cpu_baseline / --impl reference : the reference's own C++ processBlock (oracle/_ref; the C port when that build is
             absent) on the box's host cores, a bounded sample of the same workload.
The default math mode (JB_MATH_AUTO) is measured: a Punch / Saturator that feeds another plugin runs the C library's own
tanh / pow so that the chain's output is the reference's bit for bit; `fast_math` reports the same
render with the special-function-unit routines (inside 1e-5 of peak for > 99.9 % of clips, not for all).
torch is plumbing only (streams / events, torch.distributed for the barrier and the max over ranks); the engine is
juicy-audio-plugins_b200/libjuicy_batch.so called through its C ABI.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "juicy-audio-plugins_b200")
sys.path.insert(0, ROOT)

SAMPLE_RATE = 48000.0
BLOCK = 512
METRIC = "channel-samples/sec per plugin chain"
UNIT = "channel-samples/s"
FULL_CHAIN = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]

# name -> workload (BASELINE.json configs[i]); per_clip = (slot, parameter id, modulus): value = clip index mod modulus
WORKLOADS = {
    "C1": {"label": "configs[0]: JuicySaturator processBlock on a 10 s 48 kHz stereo sine sweep, 512-sample blocks",
           "chain": ["JuicySaturator"], "clips": 1, "samples": 480000, "synth": "sweep", "per_clip": None},
    "C2": {"label": "configs[1]: JuicyPunch -> JuicyWidth chain on 4096 stereo drum-hit clips (1 s, 48 kHz) per GPU",
           "chain": ["JuicyPunch", "JuicyWidth"], "clips": 4096, "samples": 48000, "synth": "drum", "per_clip": None},
    "C3": {"label": "configs[2]: JuicyTexture resonator bank on 16384 impulse-train streams (8192 stereo instances), material = clip mod 5",
           "chain": ["JuicyTexture"], "clips": 8192, "samples": 48000, "synth": "impulse", "per_clip": (0, "material", 5)},
    "C4": {"label": "configs[3]: JuicyInfer pre/post scoring on 65536 noise / sweep / impulse / drum clips",
           "chain": ["JuicyInfer"], "clips": 65536, "samples": 48000, "synth": "mixed", "per_clip": None},
    "C5": {"label": "configs[4]: full 7-plugin chain (Punch->Saturator->Texture->Width->Motion->Cohere->Infer), 32768 stereo clips "
                    "per GPU (one shard of 262144 clips over 8 GPUs)",
           "chain": FULL_CHAIN, "clips": 32768, "samples": 48000, "synth": "mixed", "per_clip": None},
}


def load_juicy_batch():
    spec = importlib.util.spec_from_file_location("juicy_batch", os.path.join(PKG, "juicy_batch.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["juicy_batch"] = mod
    spec.loader.exec_module(mod)
    return mod


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU before any pinned host memory is allocated, so
    the staging buffers of jb_process_host sit on the GPU's own NUMA node (matters once several ranks share a host)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
        return allowed
    except Exception:
        return None  # plumbing only: without NVML the rank simply stays where the launcher put it


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=100):
        self.index = index
        self.interval_ms = interval_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(w, name):
    """The `config` object of the JSON line: identical for the engine arm and the reference arm."""
    return {"workload": w["label"], "name": name, "chain": w["chain"], "clips_per_gpu": w["clips"], "channels": 2,
            "samples_per_clip": w["samples"], "sample_rate": SAMPLE_RATE, "block_size": BLOCK,
            "parameters": "plugin defaults (program 0)" + ("; %s = clip mod %d" % (w["per_clip"][1], w["per_clip"][2]) if w["per_clip"] else ""),
            "input": w["synth"],
            "cache": "inputs (%.2f GB per GPU) are larger than the 126 MB L2; no flush needed"
                     % (w["clips"] * 2 * w["samples"] * 4 / 1e9) if w["clips"] * w["samples"] * 8 > 4e8 else
                     "a 1.2 GB buffer is overwritten between timed renders (L2 flush): the clip itself fits the L2",
            "sharding": "independent clips per rank, no data-path collective; one ncclAllGather of per-clip records per step"}


def algorithmic_bytes(w):
    """8 B per channel-sample per chain (4 read + 4 written; Infer with trim = 0 dB is read-only: 4)
    plus one 64 B record per (clip, block, plugin)."""
    ch_samples = w["clips"] * 2 * w["samples"]
    n_blocks = (w["samples"] + BLOCK - 1) // BLOCK
    per = 4.0 if w["chain"] == ["JuicyInfer"] else 8.0
    return per * ch_samples + 64.0 * w["clips"] * n_blocks * len(w["chain"])


# ---------------------------------------------------------------------------------------------- CPU reference

def cpu_worker(spec):
    """One host process of the CPU arm: its slice of the clips through the reference chain.  Inputs come from the numpy
    generator (oracle/synth_np.py): this process never loads the product library.
    (Threads of one process share a core in some sandboxes, so the CPU arm uses processes.)"""
    from oracle import refhost, port, synth_np
    cls = refhost.RefPlugin if spec["kind"] == "reference" else port.PortPlugin
    lo, hi, n = spec["lo"], spec["hi"], spec["samples"]
    chain, per_clip = spec["chain"], spec.get("per_clip")
    clips = synth_np.synth_clips(spec["synth"], lo, hi - lo, n, 2, SAMPLE_RATE)
    work = np.empty_like(clips)
    rec = np.zeros((hi - lo, 16), dtype=np.float32)
    # one instance per (plugin, parameter value): per-clip values render as groups of clips, like the engine's parameter sets
    variants = range(per_clip[2]) if per_clip else [None]
    plugins = {}
    for v in variants:
        ps = [cls(p, 2, SAMPLE_RATE, BLOCK) for p in chain]
        if v is not None:
            ps[per_clip[0]].set_param(per_clip[1], float(v))
        plugins[v] = ps
    groups = {v: [c for c in range(lo, hi) if (c % per_clip[2]) == v] for v in variants} if per_clip else {None: list(range(lo, hi))}

    def render():
        np.copyto(work, clips)
        for v, members in groups.items():
            for p in plugins[v]:  # plugin by plugin == block by block: every plugin is causal with identical blocking
                if per_clip:
                    for c in members:
                        p.lib.render_clips(p.h, work[c - lo].ctypes.data, 1, n, BLOCK, SAMPLE_RATE, rec[c - lo].ctypes.data)
                else:
                    p.lib.render_clips(p.h, work.ctypes.data, hi - lo, n, BLOCK, SAMPLE_RATE, rec.ctypes.data)

    for _ in range(spec["warmup"]):
        render()
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    sys.stdin.readline()
    t0 = time.time()
    for _ in range(spec["steps"]):
        render()
    t1 = time.time()
    sys.stdout.write(json.dumps({"start": t0, "end": t1}) + "\n")
    sys.stdout.flush()
    return 0


def cpu_reference_run(w, n_clips, procs, warmup, steps):
    """The reference's own processBlock (oracle/_ref; the C port if that build is absent) over n_clips clips of workload
    `w`, split over `procs` host processes.  Returns (kind, processes used, seconds for `steps` passes)."""
    from oracle import refhost, port
    if refhost.available():
        kind = "reference"
    else:
        kind = "port"
        port.lib()
    procs = max(1, min(procs, n_clips))
    per = (n_clips + procs - 1) // procs
    slices = [(t * per, min(n_clips, (t + 1) * per)) for t in range(procs) if t * per < n_clips]
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    workers = []
    for lo, hi in slices:
        spec = {"kind": kind, "lo": lo, "hi": hi, "samples": w["samples"], "warmup": warmup, "steps": steps,
                "chain": w["chain"], "synth": w["synth"], "per_clip": w["per_clip"]}
        workers.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", json.dumps(spec)],
                                        stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env))
    for wk in workers:
        line = wk.stdout.readline()
        if line.strip() != "ready":
            raise RuntimeError("cpu worker failed to start: %r" % line)
    for wk in workers:
        wk.stdin.write("go\n")
        wk.stdin.flush()
    stamps = [json.loads(wk.stdout.readline()) for wk in workers]
    for wk in workers:
        wk.wait()
    seconds = max(s["end"] for s in stamps) - min(s["start"] for s in stamps)
    return kind, len(slices), seconds


def cpu_sample_clips(w, threads, override=0):
    """Bounded sample of the workload for the CPU legs: ~4 M channel-samples per host thread and pass, at least one clip
    per thread where the workload has that many (C1 is a single clip: one process renders it)."""
    if override:
        return min(override, w["clips"])
    per_thread = max(1, int(4.0e6 // (2 * w["samples"] * max(1, len(w["chain"]) // 2))))
    return max(1, min(w["clips"], threads * per_thread))


def cpu_baseline(w, threads, override=0, passes=2):
    n_clips = cpu_sample_clips(w, threads, override)
    kind, used, secs = cpu_reference_run(w, n_clips, threads, 1, passes)
    return {"value": passes * n_clips * 2 * w["samples"] / secs, "unit": UNIT, "cores": used, "kind": kind,
            "sample": "%d of the %d clips (x 2 ch x %d samples) through %s in %d-sample blocks, %d passes, %d host processes"
                      % (n_clips, w["clips"], w["samples"], " -> ".join(w["chain"]), BLOCK, passes, used)}


def pick_workload(args, world):
    name = args.config or "C5"
    w = dict(WORKLOADS[name])
    if args.clips:
        w["clips"] = args.clips
    if args.samples:
        w["samples"] = args.samples
    return name, w


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return 0
    name, w = pick_workload(args, max(world, args.gpus))
    threads = host_threads()
    n_clips = cpu_sample_clips(w, threads, args.cpu_clips)
    kind, used, total = cpu_reference_run(w, n_clips, threads, args.warmup, args.steps)
    value = n_clips * 2 * w["samples"] * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(w, name),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": "each step renders a bounded sample of %d of the workload's %d clips (x 2 ch x %d samples) through %s "
                                   "in %d-sample blocks on %d host processes" % (n_clips, w["clips"], w["samples"], " -> ".join(w["chain"]), BLOCK, used)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- engine arm

class Bench:
    """One workload on this rank's GPU: device-resident steps, end-to-end steps, clocks."""

    def __init__(self, jb, torch, w, name, local, rank, world, stream, comm_id=None):
        self.jb, self.torch, self.w, self.name = jb, torch, w, name
        self.local, self.rank, self.world, self.stream = local, rank, world, stream
        n_clips, n = w["clips"], w["samples"]
        self.count = n_clips * 2 * n
        self.d_in = torch.empty(self.count, dtype=torch.float32, device="cuda")
        # JuicyInfer with trim = 0 dB leaves the audio untouched (JuicyInfer/PluginProcessor.cpp:79: applyGain(1) is a no-op):
        # it scores in place, as processBlock does; every other workload renders out of place
        self.in_place = w["chain"] == ["JuicyInfer"]
        self.d_out = self.d_in if self.in_place else torch.empty(self.count, dtype=torch.float32, device="cuda")
        jb.synth_fill_device(self.d_in.data_ptr(), w["synth"], rank * n_clips, n_clips, 2, n, SAMPLE_RATE, device=local,
                             stream=stream.cuda_stream)
        self.eng = jb.BatchProcessor(w["chain"], n_clips, device=local)
        self.eng.set_stream(stream.cuda_stream)
        if w["per_clip"]:
            slot, pid, mod = w["per_clip"]
            for c in range(n_clips):
                self.eng.setParameterClips(pid, float((rank * n_clips + c) % mod), c, 1, slot)
        self.eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        self.last_slot = len(w["chain"]) - 1
        self.d_gather = None
        if world > 1:
            L = jb.lib()
            jb._check(L.jb_comm_init_rank(self.eng._h, comm_id, world, rank))
            pitch = L.jb_record_pitch(self.eng._h)
            self.d_gather = torch.empty(world * 16 * pitch, dtype=torch.float32, device="cuda")
        # a single clip fits the L2: overwrite a buffer larger than the L2 between timed renders
        self.flush = torch.empty(300 * 1024 * 1024, dtype=torch.float32, device="cuda") if self.count * 8 <= 4e8 else None

    def step(self):
        self.eng.reset()
        if self.flush is not None:
            self.flush.zero_()
        self.eng.process_device(self.d_in.data_ptr(), self.d_out.data_ptr(), self.w["samples"])
        if self.world > 1:
            # per-clip records -> every rank (north_star: NCCL only to gather per-clip scores), issued by the library
            self.jb._check(self.jb.lib().jb_gather_records(self.eng._h, self.last_slot, ctypes.c_void_p(self.d_gather.data_ptr())))

    def timed(self, steps, warmup, barrier):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.step()
            torch.cuda.synchronize()
            barrier()
            self.eng.kernel_time_ms()
            self.eng.enable_slot_timing(True)
            launches0 = self.jb.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(self.stream)
            for _ in range(steps):
                self.step()
            e1.record(self.stream)
            torch.cuda.synchronize()
            barrier()
            ms_total = e0.elapsed_time(e1)
            kernel_ms, kernel_renders = self.eng.kernel_time_ms()
            self.slot_times = self.eng.slot_times_ms()
            self.eng.enable_slot_timing(False)
            launches = self.jb.launch_count() - launches0
            coop, lane = self.eng.path_launches()
        if self.flush is not None:
            # the flush is inside the event bracket: report the renders alone (library events around its launches)
            ms_total = kernel_ms
        return ms_total, kernel_ms, kernel_renders, launches, coop, lane

    def e2e(self, steps, barrier):
        """jb_process_host with pinned HOST buffers: H2D of the inputs and D2H of the results inside the timed region."""
        jb, torch = self.jb, self.torch
        n_clips, n = self.w["clips"], self.w["samples"]
        h_in = jb.PinnedBuffer((n_clips, 2, n))
        # big batches render in place on the host side too (h_out == h_in halves the pinned memory: 8 ranks x 12.6 GB)
        in_place = self.count * 4 > 4e9
        h_out = h_in if in_place else jb.PinnedBuffer((n_clips, 2, n))
        jb._check(jb.lib().jb_copy_to_host(self.local, h_in.array.ctypes.data, self.d_in.data_ptr(), self.count * 4))
        rec = None

        def one():
            self.eng.reset()
            self.eng.process_host_ptr(h_in.array.ctypes.data, h_out.array.ctypes.data, n)
            return self.eng.getLatestMetrics(self.last_slot)

        one()
        if in_place:  # the warm-up overwrote the input: restore it, so the timed steps render the same audio
            jb._check(jb.lib().jb_copy_to_host(self.local, h_in.array.ctypes.data, self.d_in.data_ptr(), self.count * 4))
        torch.cuda.synchronize()
        per_step = []
        for _ in range(steps):   # every step timed on its own (barrier + synchronize on both sides)
            barrier()
            t0 = time.perf_counter()
            rec = one()
            torch.cuda.synchronize()
            per_step.append(time.perf_counter() - t0)
        barrier()
        # The floor under those steps on THIS box, now: the same bytes as two plain pinned copies, host -> device and
        # device -> host at once (what jb_process_host would take if rendering cost nothing and needed no pipeline ramp).
        floor = None
        try:
            t_in, t_out = torch.from_numpy(h_in.array.reshape(-1)), torch.from_numpy(h_out.array.reshape(-1))
            s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
            best = []
            for _ in range(3):
                torch.cuda.synchronize()
                barrier()
                t0 = time.perf_counter()
                with torch.cuda.stream(s_up):
                    self.d_out.copy_(t_in, non_blocking=True)
                if not (in_place and self.w["chain"] == ["JuicyInfer"]):   # a scoring run in place brings no audio back
                    with torch.cuda.stream(s_down):
                        t_out.copy_(self.d_in, non_blocking=True)
                torch.cuda.synchronize()
                best.append(time.perf_counter() - t0)
            floor = min(best[1:])
            barrier()
        except Exception:
            floor = None
        self.pcie_floor_s = floor
        h_in.free()
        if h_out is not h_in:
            h_out.free()
        return per_step, rec, in_place

    def e2e_pcm16(self, steps, barrier):
        """The same through jb_process_host_pcm16: 16-bit PCM host buffers (what audio files hold), converted on the device;
        half the bytes per direction.  In place on the host side."""
        jb, torch = self.jb, self.torch
        n_clips, n = self.w["clips"], self.w["samples"]
        h = jb.PinnedBuffer((n_clips, n))          # float32 words = 2 x int16: [n_clips][2 ch][n] int16
        pcm = h.array.view(np.int16).reshape(n_clips, 2, n)
        tmp = np.empty((min(n_clips, 512), 2, n), dtype=np.float32)

        def fill():
            for c0 in range(0, n_clips, tmp.shape[0]):
                k = min(tmp.shape[0], n_clips - c0)
                jb._check(jb.lib().jb_copy_to_host(self.local, tmp.ctypes.data, self.d_in.data_ptr() + c0 * 2 * n * 4, k * 2 * n * 4))
                np.clip(np.rint(tmp[:k] * 32768.0), -32767, 32767, out=tmp[:k])
                pcm[c0:c0 + k] = tmp[:k]

        fill()
        self.eng.reset()
        self.eng.process_host_pcm16_ptr(pcm.ctypes.data, pcm.ctypes.data, n)
        fill()
        torch.cuda.synchronize()
        per_step = []
        for _ in range(steps):
            barrier()
            t0 = time.perf_counter()
            self.eng.reset()
            self.eng.process_host_pcm16_ptr(pcm.ctypes.data, pcm.ctypes.data, n)
            self.eng.getLatestMetrics(self.last_slot)
            torch.cuda.synchronize()
            per_step.append(time.perf_counter() - t0)
        barrier()
        h.free()
        return per_step

    def close(self):
        try:
            self.eng.close()
        except Exception:
            pass
        self.d_in = self.d_out = self.d_gather = self.flush = None
        self.torch.cuda.empty_cache()


def measure(jb, torch, dist, w, name, local, rank, world, stream, steps, warmup, e2e_steps, comm_id, peak, peak_src,
            with_fast=False, with_pcm=False):
    """Device-resident + end-to-end measurement of one workload; returns the result dict on rank 0 (None elsewhere)."""
    def barrier():
        if world > 1:
            dist.barrier()

    b = Bench(jb, torch, w, name, local, rank, world, stream, comm_id)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, kernel_ms, kernel_renders, launches, coop, lane = b.timed(steps, warmup, barrier)
    slot_times = list(b.slot_times)   # of the default math mode (the fast-math leg below times its own)
    # The end-to-end steps are polled once a second instead of ten times: with the 100 ms poll running, the same steps took
    # 348 / 372 ms (median of 7) against 321 / 281 ms without it (profiles/r02_e2e_smi_poll.txt; every query takes driver
    # locks that the pipeline's ~1000 copies, launches and event records per step also need).  JB_BENCH_SMI_E2E=0: no poll.
    clocks = sampler.stop() if rank == 0 else None
    slow = ClockSampler(local, 1000)
    if rank == 0 and os.environ.get("JB_BENCH_SMI_E2E", "1") != "0":
        slow.start()
    e2e_steps_s, rec_host, e2e_in_place = b.e2e(e2e_steps, barrier)
    pcm_steps_s = b.e2e_pcm16(e2e_steps, barrier) if with_pcm else [0.0] * e2e_steps
    if rank == 0 and slow.proc is not None:
        c2 = slow.stop()
        clocks["e2e_region"] = {"sm_mhz": c2.get("sm_mhz"), "samples": c2.get("samples", 0), "reasons": c2.get("reasons", []), "interval_ms": 1000}
        clocks["reasons"] = sorted(set(clocks.get("reasons", [])) | (set(c2.get("reasons", [])) - {"no samples"}))
    fast = None
    if with_fast and world == 1:
        b.eng.set_math_mode("fast")
        f_total, f_kernel, f_renders, _, _, _ = b.timed(max(2, steps // 2), 1, barrier)
        b.eng.set_math_mode("auto")
        fast = {"ms_per_step": f_total / max(2, steps // 2), "mean_render_ms": f_kernel / max(f_renders, 1),
                "per_plugin_ms": [ms / max(k, 1) for ms, k in b.slot_times]}
    b_in_place = b.in_place
    floor_s = getattr(b, "pcie_floor_s", None)
    times = torch.tensor([ms_total, kernel_ms] + [x * 1000.0 for x in e2e_steps_s] + [x * 1000.0 for x in pcm_steps_s]
                         + [1000.0 * floor_s if floor_s else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)   # every step: the slowest rank
    tl = [float(x) for x in times.tolist()]
    ms_total, kernel_ms = tl[0], tl[1]
    e2e_list, pcm_list = tl[2:2 + e2e_steps], tl[2 + e2e_steps:2 + 2 * e2e_steps]
    floor_ms = tl[2 + 2 * e2e_steps] or None
    # The host <-> device path of these boxes is shared with other tenants (single steps scatter by tens of percent,
    # profiles/r02_e2e_geometry.txt): the end-to-end figure is taken from the MEDIAN step; mean, min and max stand beside it.
    e2e_ms = float(np.median(e2e_list)) * e2e_steps
    pcm_ms = float(np.median(pcm_list)) * e2e_steps
    b.close()
    if rank != 0:
        return None
    ch_samples_rank = w["clips"] * 2 * w["samples"]
    value = world * ch_samples_rank * steps / (ms_total / 1000.0)
    e2e_value = world * ch_samples_rank * e2e_steps / (e2e_ms / 1000.0)
    alg = algorithmic_bytes(w)
    mean_render_ms = kernel_ms / max(kernel_renders, 1)
    achieved = alg / (mean_render_ms / 1000.0) / 1e9
    kernels_per_render = launches / max(kernel_renders, 1)
    if coop >= lane and coop > 0:
        kernel_name = "jb_coop_kernel"
    elif len(w["chain"]) == 1 and not w["per_clip"]:
        kernel_name = "jb_single_kernel / jb_pair_kernel (%s)" % w["chain"][0]
    else:
        kernel_name = "render = %.0f kernel launches (one per plugin / parameter set)" % kernels_per_render
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            entry = tj.get(name) if isinstance(tj.get(name), dict) else (tj if tj.get("kernel") == kernel_name else None)
            traffic = entry.get("bytes_per_launch") if entry else None
        except Exception:
            traffic = None
    count_bytes = ch_samples_rank * 4
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "algorithmic_bytes_per_launch": alg,
                "mean_launch_ms": mean_render_ms, "launches_timed": kernel_renders, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    if len(w["chain"]) > 1 and all(n > 0 for _, n in slot_times):
        # the chain rendered as one launch per plugin: the roofline object is the DOMINANT kernel's (rank 0's events between
        # the plugins' launches); the whole render -- all launches against the chain's algorithmic bytes -- goes in `step`
        n_blocks = (w["samples"] + BLOCK - 1) // BLOCK
        per = [{"plugin": p, "mean_launch_ms": ms / n, "launches_timed": n,
                # every plugin of a chain reads and rewrites the batch in place; JuicyInfer (trim = 0 dB) only reads it
                "algorithmic_bytes_per_launch": (4.0 if p == "JuicyInfer" else 8.0) * ch_samples_rank + 64.0 * w["clips"] * n_blocks}
               for p, (ms, n) in zip(w["chain"], slot_times)]
        for k in per:
            k["achieved"] = k["algorithmic_bytes_per_launch"] / (k["mean_launch_ms"] / 1000.0) / 1e9
            k["frac"] = k["achieved"] / peak
        top = max(per, key=lambda k: k["mean_launch_ms"])
        step = {"frac": achieved / peak, "achieved": achieved, "algorithmic_bytes": alg, "ms": mean_render_ms,
                "launches": kernels_per_render, "share_of_step": top["mean_launch_ms"] / mean_render_ms}
        math_kind = "exact tanhf / powf" if top["plugin"] in ("JuicyPunch", "JuicySaturator") else "default"
        roofline = {"bound": "hbm", "achieved": top["achieved"], "peak": peak, "unit": "GB/s", "frac": top["frac"],
                    "traffic": traffic, "kernel": "%s launch of the chain (jb_pair_kernel / jb_single_kernel, %s)" % (top["plugin"], math_kind),
                    "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"], "mean_launch_ms": top["mean_launch_ms"],
                    "launches_timed": top["launches_timed"], "peak_source": peak_src,
                    "frac_of_nominal_8TBs": top["achieved"] / 8000.0, "step": step, "per_plugin": per}
    res = {
        "value": value, "unit": UNIT, "ms_per_step": ms_total / steps, "steps": steps,
        "roofline": roofline,
        # JuicyInfer (trim = 0 dB) scores without changing the audio: in place on the host only the records come back
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": count_bytes,
                "d2h_bytes_per_step": (0 if (e2e_in_place and w["chain"] == ["JuicyInfer"]) else count_bytes) + 64 * w["clips"],
                "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "host_in_place": e2e_in_place,
                "statistic": "median of the individually timed steps (max over ranks per step)",
                "ms_per_step_mean": float(np.mean(e2e_list)), "ms_per_step_min": float(min(e2e_list)), "ms_per_step_max": float(max(e2e_list)),
                "pcie_floor_ms": floor_ms, "pcie_floor_note": "the same bytes as two plain pinned copies (H2D and D2H at once, no render), "
                                                              "measured in this process right after the steps; best of 2, max over ranks",
                "api": "jb_process_host + jb_get_metrics (pinned host buffers)"},
        "gpu_launches": int(launches), "clocks": clocks, "device_in_place": b_in_place,
        "math": "auto (exact tanh / pow where a Punch / Saturator feeds another plugin)",
        "mean_juiciness": float(np.mean(rec_host[:, 13])) if rec_host is not None else None,
    }
    if with_pcm:
        res["e2e_pcm16"] = {"value": world * ch_samples_rank * e2e_steps / (pcm_ms / 1000.0), "unit": UNIT, "ms_per_step": pcm_ms / e2e_steps,
                            "ms_per_step_mean": float(np.mean(pcm_list)), "ms_per_step_min": float(min(pcm_list)),
                            "h2d_bytes_per_step": count_bytes // 2, "d2h_bytes_per_step": count_bytes // 2 + 64 * w["clips"],
                            "api": "jb_process_host_pcm16 + jb_get_metrics (pinned 16-bit PCM host buffers, device-side conversion)"}
    if fast:
        res["fast_math"] = {"ms_per_step": fast["ms_per_step"], "value": ch_samples_rank / (fast["ms_per_step"] / 1000.0),
                            "mean_render_ms": fast["mean_render_ms"],
                            "per_plugin_ms": dict(zip(w["chain"], fast["per_plugin_ms"])) if len(w["chain"]) > 1 else None,
                            "frac": alg / (fast["mean_render_ms"] / 1000.0) / 1e9 / peak,
                            "note": "jb_set_math_mode(JB_MATH_FAST): MUFU-based tanh / pow, not decision-safe for every clip"}
    return res


def run_engine_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    all_cpus = bind_to_gpu_numa_node(local)
    stdout_fd = None
    jb = load_juicy_batch()
    comm_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version (the image sets NCCL_DEBUG=VERSION) and any debug lines on fd 1; the contract is ONE JSON
        # line on stdout, so everything goes to stderr until that line is written
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the library's own communicator for the score gather: rank 0 makes the id, torch.distributed carries the 128 bytes
        idbuf = ctypes.create_string_buffer(128)
        if rank == 0:
            jb._check(jb.lib().jb_comm_unique_id(idbuf))
        t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        comm_id = bytes(t.cpu().tolist())

    name, w = pick_workload(args, world)
    stream = torch.cuda.Stream()
    peak, peak_src = measured_peak_gbs()
    e2e_steps = max(1, min(args.steps, 7 if w["clips"] * w["samples"] > 1e9 else args.steps))
    res = measure(jb, torch, dist, w, name, local, rank, world, stream, args.steps, args.warmup, e2e_steps, comm_id, peak, peak_src,
                  with_fast=True, with_pcm=True)

    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(w, name)}
        for k in ("roofline", "e2e", "e2e_pcm16", "gpu_launches", "clocks", "math", "mean_juiciness", "fast_math"):
            if k in res:
                line[k] = res[k]
        if world == 1 and not args.no_cpu:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)  # the CPU leg uses every host core again
            line["cpu_baseline"] = cpu_baseline(w, host_threads(), args.cpu_clips)
        if world == 1 and not args.no_survey and not args.config:
            # the other BASELINE.json configurations on this GPU, each measured like the headline (fewer steps)
            others = []
            for other in ("C2", "C1", "C3", "C4"):
                ow = dict(WORKLOADS[other])
                try:
                    o_steps = 10 if other == "C2" else 3
                    o_e2e = {"C2": 9, "C1": 5, "C3": 5, "C4": 3}[other]
                    r = measure(jb, torch, dist, ow, other, local, 0, 1, stream, o_steps, 3, o_e2e, None, peak, peak_src,
                                with_fast=(other == "C2"), with_pcm=(other == "C2"))
                    r["config"] = workload_config(ow, other)
                    if not args.no_cpu:
                        r["cpu_baseline"] = cpu_baseline(ow, host_threads(), 0)
                        r["e2e_vs_cpu"] = r["e2e"]["value"] / r["cpu_baseline"]["value"]
                    others.append(r)
                except Exception as exc:  # a survey line must never take the bench line down
                    others.append({"config": workload_config(ow, other), "error": repr(exc)})
            line["other_configs"] = others
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("engine", "reference"), default="engine")
    ap.add_argument("--config", choices=sorted(WORKLOADS), default=None,
                    help="workload (default: C5 = BASELINE.json configs[4], one 32768-clip shard per GPU, at every N)")
    ap.add_argument("--clips", type=int, default=0, help="override the clips per GPU of the workload")
    ap.add_argument("--samples", type=int, default=0, help="override the samples per clip of the workload")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the CPU sample (default: ~4 M channel-samples per host thread)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-survey", action="store_true", help="skip the other_configs survey (N = 1 only)")
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        return cpu_worker(json.loads(args.cpu_worker))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_engine_arm(args)


if __name__ == "__main__":
    sys.exit(main())
