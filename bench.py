#!/usr/bin/env python
"""bench.py -- channel-samples/s of the JuicySuite hot path on B200 (BASELINE.json metric).

Workload at every N: configs[1] of BASELINE.json per GPU -- the JuicyPunch -> JuicyWidth chain on
4096 stereo drum-hit clips (1 s, 48 kHz, 512-sample blocks), weak scaling (each rank renders its
own 4096 clips; clips are independent, so the data path has no collective; rank 0 gathers the
per-clip Juiciness records with one NCCL all_gather, inside the timed region).

A step = prepareToPlay-reset + one render of the whole batch, out of place (input buffer is
1.57 GB, far larger than the 126 MB L2, so no step sees a warm cache).
  value    : device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e      : the same render through jb_process_host with pinned HOST buffers, H2D/D2H inside
  roofline : algorithmic bytes per launch / mean device duration of the render kernel (events
             recorded by the library around every launch on its stream) vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the reference's own C++ processBlock (oracle/_ref, or the C port
             when the reference build is absent) on the box's host cores, bounded sample.
torch is plumbing only (device buffers, streams/events, torch.distributed); the engine is
juicy-audio-plugins_b200/libjuicy_batch.so called through its C ABI.
"""
import argparse
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

CHAIN = ["JuicyPunch", "JuicyWidth"]
SAMPLE_RATE = 48000.0
BLOCK = 512
METRIC = "channel-samples/sec per plugin chain"
UNIT = "channel-samples/s"


def load_juicy_batch():
    spec = importlib.util.spec_from_file_location("juicy_batch", os.path.join(PKG, "juicy_batch.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["juicy_batch"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_sharding():
    spec = importlib.util.spec_from_file_location("jb_sharding", os.path.join(PKG, "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
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

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


# ---------------------------------------------------------------------------------------------- CPU reference

def cpu_worker(spec):
    """One host process of the CPU arm: its slice of the clips through the reference chain.
    (Threads of one process share a core in some sandboxes, so the CPU arm uses processes.)"""
    jb = load_juicy_batch()
    from oracle import refhost, port
    cls = refhost.RefPlugin if spec["kind"] == "reference" else port.PortPlugin
    lo, hi, n = spec["lo"], spec["hi"], spec["samples"]
    clips = jb.synth_clips("drum", lo, hi - lo, n, 2, SAMPLE_RATE)
    work = np.empty_like(clips)
    rec = np.zeros((hi - lo, 16), dtype=np.float32)
    plugins = [cls(p, 2, SAMPLE_RATE, BLOCK) for p in CHAIN]

    def render():
        np.copyto(work, clips)
        for p in plugins:  # plugin by plugin == block by block: every plugin is causal with identical blocking
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


def cpu_reference_run(n_clips, n_samples, procs, warmup, steps):
    """The reference's own processBlock (oracle/_ref; the C port if that build is absent) over n_clips
    drum-hit clips, split over `procs` host processes.  Returns (kind, processes used, seconds for `steps` passes)."""
    from oracle import refhost, port
    if refhost.available():
        kind = "reference"
    else:
        kind = "port"
        port.lib()
    per = (n_clips + procs - 1) // procs
    slices = [(t * per, min(n_clips, (t + 1) * per)) for t in range(procs) if t * per < n_clips]
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    workers = []
    for lo, hi in slices:
        spec = {"kind": kind, "lo": lo, "hi": hi, "samples": n_samples, "warmup": warmup, "steps": steps}
        workers.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", json.dumps(spec)],
                                        stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env))
    for w in workers:
        line = w.stdout.readline()
        if line.strip() != "ready":
            raise RuntimeError("cpu worker failed to start: %r" % line)
    for w in workers:
        w.stdin.write("go\n")
        w.stdin.flush()
    stamps = [json.loads(w.stdout.readline()) for w in workers]
    for w in workers:
        w.wait()
    seconds = max(s["end"] for s in stamps) - min(s["start"] for s in stamps)
    return kind, len(slices), seconds


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    jb = load_juicy_batch()
    threads = host_threads()
    n_clips = args.cpu_clips or max(threads * 32, 64)
    kind, used, total = cpu_reference_run(n_clips, args.samples, threads, args.warmup, args.steps)
    ch_samples = n_clips * 2 * args.samples
    value = ch_samples * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_clips_note="reference arm renders a bounded sample of %d clips per step" % n_clips),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": "%d drum-hit clips x 2 ch x %d samples per step, %s -> %s processBlock in %d-sample blocks, "
                                   "%d host processes" % (n_clips, args.samples, CHAIN[0], CHAIN[1], BLOCK, used)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_clips_note=None):
    cfg = {"workload": "configs[1]: JuicyPunch -> JuicyWidth chain on %d stereo drum-hit clips (1 s, 48 kHz) per GPU"
                       % args.clips,
           "chain": CHAIN, "clips_per_gpu": args.clips, "channels": 2, "samples_per_clip": args.samples,
           "sample_rate": SAMPLE_RATE, "block_size": BLOCK, "parameters": "plugin defaults (program 0)",
           "cache": "inputs (%.2f GB per GPU) are larger than the 126 MB L2; no flush needed"
                    % (args.clips * 2 * args.samples * 4 / 1e9),
           "sharding": "independent clips per rank, no data-path collective; one all_gather of per-clip records"}
    if n_clips_note:
        cfg["note"] = n_clips_note
    return cfg


# ---------------------------------------------------------------------------------------------- engine arm

FULL_CHAIN = ["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion", "JuicyCohere", "JuicyInfer"]


def survey_other_configs(jb, device, peak_gbs):
    """Device-resident renders of BASELINE.json configs[0], [2], [3] and one 8-GPU shard of [4] (SURVEY.md §8(d) shapes),
    timed with the library's CUDA events around its kernel launches; 1 warm-up + 2 timed renders each."""
    specs = [
        ("configs[0] C1: JuicySaturator, 1 clip x 10 s sine sweep", ["JuicySaturator"], 1, 480000, "sweep", None),
        ("configs[2] C3: JuicyTexture, 8192 stereo instances (16384 streams), impulse trains, material = clip mod 5",
         ["JuicyTexture"], 8192, 48000, "impulse", ("material", 5)),
        ("configs[3] C4: JuicyInfer scoring, 65536 clips (noise / sweep / impulse / drum mix)", ["JuicyInfer"], 65536, 48000, "mixed", None),
        ("configs[4] C5: full 7-plugin chain, one GPU's shard of 262144 clips (32768)", FULL_CHAIN, 32768, 48000, "mixed", None),
    ]
    out = []
    for name, chain, n_clips, n, synth, clipmod in specs:
        try:
            buf = jb.DeviceBuffer(n_clips * 2 * n * 4, device)
            # the big configurations render in place (C4 is 25 GB of input); C3's five concurrent launches go out of place
            dst = jb.DeviceBuffer(n_clips * 2 * n * 4, device) if clipmod else buf
            jb.synth_fill_device(buf.ptr.value, synth, 0, n_clips, 2, n, SAMPLE_RATE, device=device, stream=0)
            eng = jb.BatchProcessor(chain, n_clips, device=device)
            if clipmod:
                pid, k = clipmod
                for c in range(n_clips):
                    eng.setParameterClips(pid, float(c % k), c, 1, 0)
            eng.prepareToPlay(SAMPLE_RATE, BLOCK)
            eng.reset()
            eng.process_device(buf.ptr.value, dst.ptr.value, n)
            eng.synchronize()
            eng.kernel_time_ms()
            steps = 2
            for _ in range(steps):
                eng.reset()
                eng.process_device(buf.ptr.value, dst.ptr.value, n)
            ms, launches = eng.kernel_time_ms()
            ms /= steps
            ch_samples = n_clips * 2 * n
            read_only = chain == ["JuicyInfer"]
            n_blocks = (n + BLOCK - 1) // BLOCK
            alg = (4.0 if read_only else 8.0) * ch_samples + 64.0 * n_clips * n_blocks * len(chain)
            out.append({"workload": name, "chain": chain, "clips": n_clips, "samples_per_clip": n, "ms_per_render": ms,
                        "launches_per_render": launches / steps, "value": ch_samples / (ms / 1000.0), "unit": UNIT,
                        "in_place": dst is buf, "algorithmic_bytes": alg, "frac_of_hbm_peak": alg / (ms / 1000.0) / 1e9 / peak_gbs,
                        "math": "auto (fast: no resonant Texture material behind a shaper in these configurations)"})
            eng.close()
            if dst is not buf:
                dst.free()
            buf.free()
        except Exception as exc:  # a survey line must never take the bench line down
            out.append({"workload": name, "error": str(exc)})
    return out


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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version (the image sets NCCL_DEBUG=VERSION) and any debug lines on fd 1; the contract is ONE JSON
        # line on stdout, so everything goes to stderr until that line is written
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    jb = load_juicy_batch()
    sharding = load_sharding()
    n_clips, n = args.clips, args.samples
    count = n_clips * 2 * n
    d_in = torch.empty(count, dtype=torch.float32, device="cuda")
    d_out = torch.empty(count, dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    jb.synth_fill_device(d_in.data_ptr(), "drum", rank * n_clips, n_clips, 2, n, SAMPLE_RATE, device=local,
                         stream=stream.cuda_stream)
    eng = jb.BatchProcessor(CHAIN, n_clips, device=local)
    eng.set_stream(stream.cuda_stream)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    last_slot = len(CHAIN) - 1
    rec_bytes = 16 * 4 * n_clips
    # device view of the engine's SoA metrics record [16][clipPitch] for the gather
    pitch = sharding.clip_pitch(n_clips)

    class _DeviceView:  # zero-copy torch view of the engine's record block (plumbing for the NCCL gather)
        __cuda_array_interface__ = {"shape": (16 * pitch,), "typestr": "<f4", "version": 2,
                                    "data": (eng.metrics_device_ptr(last_slot), False)}

    local_rec = torch.as_tensor(_DeviceView(), device="cuda")
    first_clip, _ = sharding.shard_range(world * n_clips, rank, world)  # weak scaling: n_clips per rank
    assert first_clip == rank * n_clips

    def barrier():
        if world > 1:
            dist.barrier()

    def step():
        eng.reset()
        eng.process_device(d_in.data_ptr(), d_out.data_ptr(), n)
        if world > 1:
            # per-clip records -> every rank (north_star: NCCL only to gather per-clip scores)
            sharding.gather_records(local_rec, world, dist)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        barrier()
        eng.kernel_time_ms()
        launches0 = jb.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms_total = e0.elapsed_time(e1)
        kernel_ms, kernel_launches = eng.kernel_time_ms()
        launches = jb.launch_count() - launches0
        coop_launches, lane_launches = eng.path_launches()
        kernel_name = "jb_coop_kernel" if coop_launches >= lane_launches else "jb_process_kernel"

        # ---- end to end through the C ABI with pinned host buffers
        h_in = jb.PinnedBuffer((n_clips, 2, n))
        h_out = jb.PinnedBuffer((n_clips, 2, n))
        jb._check(jb.lib().jb_copy_to_host(local, h_in.array.ctypes.data, d_in.data_ptr(), count * 4))
        rec_host = None

        def e2e_step():
            eng.reset()
            eng.process_host_ptr(h_in.array.ctypes.data, h_out.array.ctypes.data, n)
            return eng.getLatestMetrics(last_slot)

        e2e_step()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rec_host = e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([ms_total, e2e_s * 1000.0, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms_max = [float(x) for x in times.tolist()]

    ch_samples_rank = n_clips * 2 * n
    value = world * ch_samples_rank * args.steps / (ms_total / 1000.0)
    e2e_value = world * ch_samples_rank * args.steps / (e2e_ms / 1000.0)

    # roofline of the render kernel: 8 B per channel-sample (4 read + 4 written) plus one 64 B
    # record per (clip, block, plugin)
    n_blocks = (n + BLOCK - 1) // BLOCK
    alg_bytes = 8.0 * ch_samples_rank + 64.0 * n_clips * n_blocks * len(CHAIN)
    peak, peak_src = measured_peak_gbs()
    mean_launch_ms = kernel_ms / max(kernel_launches, 1)
    achieved = alg_bytes / (mean_launch_ms / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("bytes_per_launch") if tj.get("kernel") == kernel_name else None
        except Exception:
            traffic = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kernel_name, "algorithmic_bytes_per_launch": alg_bytes,
                         "mean_launch_ms": mean_launch_ms, "launches_timed": kernel_launches, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": count * 4,
                    "d2h_bytes_per_step": count * 4 + rec_bytes, "ms_per_step": e2e_ms / args.steps,
                    "api": "jb_process_host + jb_get_metrics (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "mean_juiciness": float(np.mean(rec_host[:, 13])) if rec_host is not None else None,
        }
        if world == 1 and not args.no_cpu:
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)  # the CPU leg uses every host core again
            threads = host_threads()
            cpu_clips = args.cpu_clips or max(threads * 32, 64)
            kind, used, secs = cpu_reference_run(cpu_clips, n, threads, 1, 2)
            line["cpu_baseline"] = {
                "value": 2 * cpu_clips * 2 * n / secs, "unit": UNIT, "cores": used, "kind": kind,
                "sample": "%d of the %d drum-hit clips (x 2 ch x %d samples), 2 passes, %d host processes"
                          % (cpu_clips, n_clips, n, used)}
        if world == 1 and not args.no_survey:
            # the other BASELINE.json configurations, device resident, one GPU (explanatory: the bench line above is
            # configs[1]); freed first: C4 alone holds 25 GB of input
            eng.close()
            del d_in, d_out, local_rec
            torch.cuda.empty_cache()
            line["other_configs"] = survey_other_configs(jb, local, peak)
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    try:
        eng.close()
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("engine", "reference"), default="engine")
    ap.add_argument("--clips", type=int, default=4096, help="clips per GPU (BASELINE.json configs[1]: 4096)")
    ap.add_argument("--samples", type=int, default=48000, help="samples per clip (1 s at 48 kHz)")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the CPU sample (default 4 per host thread)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
