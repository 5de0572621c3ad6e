#!/usr/bin/env python
"""bench.py -- output Msamples/s of the windowed-sinc resampling hot path at BASELINE.json's metric config
(stereo float32, 44.1 kHz -> 48 kHz, preset -3 = 380 filters x 380 taps, interpolated), through the C ABI of
libresampler_b200.so.

A *step* is one pass of the hot path over one batch of synthetic input: STREAMS independent stereo streams, each advanced
by LAUNCHES blocks of FRAMES input frames -- one resampleBatchProcessInterleavedDevice call (include/resampler_b200.h) per
block, all streams in that one launch.  Stream state carries over from block to block exactly as in a real conversion.
With the defaults a step is ~50 ms of GPU work and the timed region of 20 steps ~1 s: the headline is a SUSTAINED number
(clocks and power sampled inside the region only); a 20-launch burst on the cold GPU is reported beside it.

  value             sustained whole-job output samples/s, inputs resident in HBM (CUDA events, max over ranks)
  burst_value       20 launches on the cold GPU, before the sustained region (round 1 reported a burst)
  strict_fp32_value the same workload with the tensor-core kernel switched off (FFMA kernels: fp32 accuracy relative to
                    every 380-tap window; the tensor-core form is exact to 2^-24 of a 0.43 s block's peak)
  e2e               the same metric through the reference-facing host-pointer API (resampleProcessInterleaved,
                    include/resampler.h) with pinned HOST buffers: H2D of every block's input and D2H of its output inside
                    the timed region
  roofline          the convolution kernel alone: algorithmic bytes 4*(1+1/ratio) per output sample x samples
                    per launch / the kernel's own launch duration, measured live with CUDA events recorded by the library
                    around each launch on its stream
  parity_check      the LAST TIMED launch's output of two streams against the oracle (outside the timed region)
  configs           every BASELINE.json config and preset -1..-4 at bench size: Gsamples/s, HBM fraction, the CPU path's number
  cpu_baseline      the UNMODIFIED reference (oracle/_ref/libartref.so, kind "reference") or, when that did not travel, the
                    oracle port -- on this host's cores, bounded sample

`--impl reference` times the reference's own CPU implementation instead (rank 0 only).
Multi-GPU: independent streams are sharded across ranks with no data-path collective (weak scaling; BASELINE config 4's
1024 contexts and config 3's 64 channels are split over the ranks); torch.distributed (NCCL) only provides the barrier and
the max-over-ranks of the device time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

PRESETS = {1: (48, 48), 2: (320, 156), 3: (380, 380), 4: (988, 988)}      # (filters, taps), artest.c:154-169
FLAGS = 0x1 | 0x2                                 # SUBSAMPLE_INTERPOLATE | BLACKMAN_HARRIS (artest.c:126)
METRIC = "output Msamples/sec at preset -3 (380-tap), 44.1k->48k"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return {}, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / power / throttle reasons (B200_PROFILING.md recipe); every row is stamped on arrival so that
    summary() can keep the samples that fall INSIDE the timed region only."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int, period_ms: int = 10):
        self.device, self.rows, self.proc, self.period = device, [], None, period_ms

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period), "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            time.sleep(0.3)                     # let the first rows arrive before the region starts
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [f.strip() for f in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, r in self.rows:
            if not (t0 <= stamp <= t1):
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_median": float(np.median(pw)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm), "sampled": "inside the timed region only"}


# ------------------------------------------------------------------------------- multi-rank plumbing

def shard_streams(total_streams: int, world: int, rank: int):
    """Independent streams are the unit of sharding: contiguous, sizes differing by at most one.
    Returns (first, count) of the streams this rank owns."""
    base, extra = divmod(total_streams, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def reduce_over_ranks(dist, device, elapsed_ms: float, units: float):
    """The job's time is the slowest rank's (MAX), its work the sum over ranks: no other collective exists
    on this path.  Works on any backend (nccl on the GPU box, gloo in the CPU tests)."""
    import torch
    t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
    u = torch.tensor([units], device=device, dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


# ------------------------------------------------------------------------------------ workloads

class Workload:
    """One configuration of the path: `streams` contexts of `ch` channels, blocks of `frames` input frames per launch."""

    def __init__(self, name, ch, preset, src, dst, streams, frames, lowpass_hz=0, fixed=False, biquad=False,
                 asrc_blocks=0, ratio=None, tensor_mode=1):
        self.tensor_mode = tensor_mode                  # resampleB200SetTensorPath while this workload runs (3: also fixed-ratio contexts)
        self.name, self.ch, self.preset, self.src, self.dst = name, ch, preset, src, dst
        self.streams, self.frames, self.lowpass_hz, self.fixed, self.biquad = streams, frames, lowpass_hz, fixed, biquad
        self.asrc_blocks = asrc_blocks                  # > 0: one stream, asrc_blocks blocks of `frames` per launch, ratio swept
        self.filters, self.taps = PRESETS[preset]
        self.ratio = ratio if ratio is not None else dst / src
        self.bytes_per_output_sample = 4.0 * (1.0 + 1.0 / self.ratio)       # one write + 1/ratio reads

    def init_args(self):
        if self.fixed:
            return ("fixed", (self.ch, self.taps, self.filters, float(self.src), float(self.dst), int(self.lowpass_hz), FLAGS | 0x4))
        return ("init", (self.ch, self.taps, self.filters, self.lowpass_hz * 2.0 / self.src, FLAGS))


METRIC_WORKLOAD = dict(ch=2, preset=3, src=44100, dst=48000)


# --------------------------------------------------------------------------------------- CPU reference

def cpu_run(w: Workload, threads: int, seconds: float = 1.5, block: int = 16384, multithreaded_flag=False):
    """The reference's CPU path (oracle/_ref = the unmodified reference; else the oracle port) on `threads` host threads, one
    context per thread (BASELINE.md section 3 mode iii), art.c-sized calls of 16384 frames.  The threads exist and have made one
    call before the clock starts; each then runs a fixed number of calls sized from that first call to last ~`seconds`.
    Returns (output samples, seconds, kind)."""
    import artlibs as A
    ref = A.reference()
    kind = "reference" if ref is not None else "port"
    make = A.reference_stream if ref is not None else A.oracle_stream
    rng = np.random.default_rng(1234)
    ch = w.ch
    if w.asrc_blocks:
        block = w.frames
    x = rng.uniform(-0.5, 0.5, (block, ch)).astype(np.float32)
    flags = FLAGS | (0x8 if multithreaded_flag else 0)

    def new_stream():
        if w.fixed:
            s = make(ch, w.taps, w.filters, flags=flags | 0x4, fixed=(w.src, w.dst, w.lowpass_hz))
        else:
            s = make(ch, w.taps, w.filters, w.lowpass_hz * 2.0 / w.src, flags=flags)
        s.advance(w.taps / 2)
        return s

    streams = [new_stream() for _ in range(threads)]
    cap = int(block * w.ratio) + w.taps + 16
    start, stop = threading.Barrier(threads + 1), threading.Barrier(threads + 1)
    plan = {"calls": 0}
    made = [0] * threads
    first_call = [0.0] * threads

    def work(t):
        s = streams[t]
        out = np.empty((cap, ch), np.float32)
        xin = x.copy()
        fn = s.lib.resampleProcessInterleaved if kind == "reference" else s.lib.oracle_process_interleaved
        xp, op = xin.ctypes.data_as(A.f32p), out.ctypes.data_as(A.f32p)
        stages = None
        if w.biquad and kind == "reference":      # art.c:848-851, :1011-1017: two lowpass sections per channel ahead of a downsampler
            co = A.BiquadCoefficients()
            s.lib.biquad_lowpass(C.byref(co), 0.45 * w.dst / w.src)
            stages = [[A.Biquad() for _ in range(ch)] for _ in range(2)]
            for st in stages:
                for q in st:
                    s.lib.biquad_init(C.byref(q), C.byref(co), 1.0)

        def one(k):
            if stages:
                for st in stages:
                    for c in range(ch):
                        s.lib.biquad_apply_buffer(C.byref(st[c]), C.cast(C.addressof(xp.contents) + 4 * c, A.f32p), block, ch)
            r = w.ratio if not w.asrc_blocks else 1.0 + 1e-4 * math.sin(2 * math.pi * k / 64.0)
            return fn(s.ctx, xp, block, op, cap, r).output_generated

        t0 = time.perf_counter()
        one(0)
        first_call[t] = time.perf_counter() - t0
        start.wait()                              # main thread sizes the run between the two barriers
        start.wait()
        n = 0
        for k in range(plan["calls"]):
            n += one(k + 1)
        made[t] = n
        stop.wait()

    pool = [threading.Thread(target=work, args=(t,), daemon=True) for t in range(threads)]
    for th in pool:
        th.start()
    start.wait()
    plan["calls"] = max(1, int(seconds / max(first_call)))
    t0 = time.perf_counter()
    start.wait()
    stop.wait()
    dt = time.perf_counter() - t0
    for th in pool:
        th.join()
    return sum(made) * ch, dt, kind, plan["calls"]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    w = Workload("metric", streams=cores, frames=16384, **METRIC_WORKLOAD)
    for _ in range(max(0, min(args.warmup, 2))):
        cpu_run(w, cores, seconds=0.3)
    samples, secs, calls = 0, 0.0, 0
    kind = "reference"
    for _ in range(args.steps):
        s, dt, kind, calls = cpu_run(w, cores, seconds=2.5)       # long enough that the slowest thread's tail at the closing barrier is ~1 %
        samples += s; secs += dt
    value = samples / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "stereo float32 44.1k->48k preset -3 (380x380, interpolated); CPU reference, "
                               f"{cores} independent stereo streams (one per host thread), 16384-frame calls; same resampling "
                               "configuration as the GPU arm (stream count and call size differ: a CPU samples/s figure does not depend on them)"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} streams x {calls} calls of 16384 input frames per step (threads started before the clock)"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm

class DeviceBatch:
    """Contexts of one workload, a ring of device input/output blocks larger than L2 together, and the launch call."""

    def __init__(self, lib, pkg, torch, dev, w: Workload, ring: int, seed: int, stream):
        self.lib, self.pkg, self.torch, self.w, self.ring = lib, pkg, torch, w, ring
        kind, a = w.init_args()
        n = w.streams
        self.ctxs = [(lib.resampleFixedRatioInit if kind == "fixed" else lib.resampleInit)(*a) for _ in range(n)]
        assert all(self.ctxs), "resampleInit failed"
        for c in self.ctxs:
            lib.resampleAdvancePosition(c, w.taps / 2)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        frames_total = w.frames * max(1, w.asrc_blocks)
        # inputs: uniform in [-0.5, 0.5) (artest.c:744-754 style), resident in HBM
        self.x = torch.rand((ring, n, frames_total, w.ch), device=dev, dtype=torch.float32, generator=gen) - 0.5
        self.cap = int(frames_total * max(w.ratio, 1.0 + 2e-4 if w.asrc_blocks else w.ratio)) + w.taps + 16 + 8 * max(1, w.asrc_blocks)
        self.y = torch.empty((ring, n, self.cap, w.ch), device=dev, dtype=torch.float32)
        self.stream_ptr = C.c_void_p(stream.cuda_stream)
        ctx_t = C.POINTER(pkg.Resample)
        self.ctx_arr = (ctx_t * n)(*self.ctxs)
        self.in_arr = [(C.c_void_p * n)(*[self.x[r, i].data_ptr() for i in range(n)]) for r in range(ring)]
        self.out_arr = [(C.c_void_p * n)(*[self.y[r, i].data_ptr() for i in range(n)]) for r in range(ring)]
        self.nin = (C.c_int * n)(*([w.frames] * n))
        self.nout = (C.c_int * n)(*([self.cap] * n))
        self.ratios = (C.c_double * n)(*([w.ratio] * n))
        self.res = (pkg.ResampleResult * n)()
        self.at = 0
        if w.asrc_blocks:
            nb = w.asrc_blocks
            self.bf = (C.c_int * nb)(*([w.frames] * nb))
            self.br = (C.c_double * nb)(*[1.0 + 1e-4 * math.sin(2 * math.pi * k / nb) for k in range(nb)])
            self.bres = (pkg.ResampleResult * nb)()
            self.bpos = (C.c_double * nb)()
        self.stages = None
        if w.biquad == "fused":                   # the same two sections folded into the context's filter bank (resampleB200AttachPrefilter)
            co = pkg.BiquadCoefficients()
            lib.biquad_lowpass(C.byref(co), 0.45 * w.dst / w.src)
            sec = (pkg.Biquad * 2)()
            for q in sec:
                lib.biquad_init(C.byref(q), C.byref(co), 1.0)
            for c in self.ctxs:
                assert lib.resampleB200AttachPrefilter(c, sec, 2) == 0
        elif w.biquad:                            # art.c:848-851: two lowpass sections per channel at 0.45 * dst/src ahead of a downsampler
            co = pkg.BiquadCoefficients()
            lib.biquad_lowpass(C.byref(co), 0.45 * w.dst / w.src)
            self.bq = [[(pkg.Biquad * w.ch)() for _ in range(2)] for _ in range(n)]
            for per_stream in self.bq:
                for st in per_stream:
                    for q in st:
                        lib.biquad_init(C.byref(q), C.byref(co), 1.0)
            self.stages = [(C.POINTER(pkg.Biquad) * 2)(*[C.cast(st, C.POINTER(pkg.Biquad)) for st in per_stream]) for per_stream in self.bq]

    def launch(self):
        """one block of every stream; returns output frames produced (summed over streams)"""
        r = self.at % self.ring
        self.at += 1
        w, lib = self.w, self.lib
        if self.stages:
            for i in range(w.streams):
                lib.biquad_apply_cascade_interleaved_device(self.stages[i], 2, w.ch, C.c_void_p(self.x[r, i].data_ptr()), w.frames, self.stream_ptr)
        if w.asrc_blocks:
            done = lib.resampleProcessBlocksInterleavedDevice(self.ctxs[0], C.c_void_p(self.x[r, 0].data_ptr()), self.bf, self.br, w.asrc_blocks,
                                                              C.c_void_p(self.y[r, 0].data_ptr()), self.cap, self.bres, self.bpos, self.stream_ptr)
            assert done == w.asrc_blocks
            return int(self._generated(self.bres))
        lib.resampleBatchProcessInterleavedDevice(self.ctx_arr, w.streams, self.in_arr[r], self.nin, self.out_arr[r], self.nout,
                                                  self.ratios, self.res, self.stream_ptr)
        return int(self._generated(self.res))

    @staticmethod
    def _generated(results):
        # a numpy view of the ResampleResult array: a Python loop over 1024 ctypes structs costs more than the launch it follows
        return np.frombuffer(results, dtype=np.uint32)[1::2].sum(dtype=np.int64)

    def close(self):
        for c in self.ctxs:
            self.lib.resampleFree(c)
        del self.x, self.y
        self.torch.cuda.empty_cache()


def time_launches(torch, batch: DeviceBatch, stream, launches: int, warm: int = 3):
    for _ in range(warm):
        batch.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    made = 0
    for _ in range(launches):
        made += batch.launch()
    e1.record(stream)
    torch.cuda.synchronize()
    return made, e0.elapsed_time(e1)


def path_counts(lib):
    g, p = C.c_ulonglong(), C.c_ulonglong()
    lib.resampleB200PathCounts(C.byref(g), C.byref(p))
    return g.value, p.value, lib.resampleB200TensorLaunches()


def kernel_name(before, after):
    d = [a - b for a, b in zip(after, before)]
    return "tensor" if d[2] else "periodic" if d[1] else "generic"


def parity_check(torch, batch: DeviceBatch, snapshots, last_ring: int, check_frames: int = 12000):
    """The last timed launch of `batch` against the oracle, for the streams in `snapshots` = {stream: (outputOffset,
    inputIndex)} taken right before that launch.  The oracle is put into the same state, given the last T frames the stream
    had consumed (the tail of the previous block) as its ring contents, and fed the first `check_frames` frames of the block."""
    import artlibs as A
    w = batch.w
    prev_ring = (last_ring - 1) % batch.ring
    worst, checked = 0.0, 0
    for s, (P, I) in snapshots.items():
        o = A.oracle_stream(w.ch, w.taps, w.filters, w.lowpass_hz * 2.0 / w.src, flags=FLAGS)
        st = C.cast(o.ctx, C.POINTER(A.OracleState)).contents       # the oracle's public head: put it into the context's state
        tail = batch.x[prev_ring, s, w.frames - w.taps:, :].cpu().numpy()                 # [T][ch]
        st.write_index, st.read_pos = int(I), float(P)
        for c in range(w.ch):
            for i in range(w.taps):
                st.ring[c * st.ring_len + I - w.taps + i] = float(tail[i, c])
        x = batch.x[last_ring, s, :check_frames, :].cpu().numpy()
        yo, used, made = o.process(x, int(check_frames * w.ratio) + w.taps, w.ratio)
        yg = batch.y[last_ring, s, :made, :].cpu().numpy()
        assert used == check_frames and made > 0
        worst = max(worst, A.peak_error(yg, yo))
        checked += made * w.ch
    return {"max_err": worst, "relative_to": "peak of the oracle's output", "tolerance": 1e-6, "ok": bool(worst <= 1e-6),
            "streams": sorted(snapshots), "samples_compared": checked,
            "what": "output of the last timed launch vs oracle/art_oracle.c started from the same (outputOffset, inputIndex, history)"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libresampler_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = entry.load_package()
    if not pkg.LIB_PATH.exists():
        entry.build()
    lib = pkg.load()
    assert lib.resampleB200SetDevice(local) == 0
    dev = torch.device("cuda", local)
    # a dedicated non-default stream: handle 0 (torch's default stream) would read as NULL = "the context's private stream"
    # in the C API, and events on the default stream would then time nothing
    work_stream = torch.cuda.Stream(device=dev)
    assert work_stream.cuda_stream != 0
    peaks, peak, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    streams, frames, per_step = args.streams, args.frames, args.launches_per_step
    w = Workload("metric", streams=streams, frames=frames, **METRIC_WORKLOAD)
    RING = 4                                      # 4 x (128 + 139) MiB of blocks: every launch reads and writes buffers that left L2 long ago
    batch = DeviceBatch(lib, pkg, torch, dev, w, RING, 20261017 + rank, work_stream)

    # ---- burst: 20 launches on the cold GPU (~5 ms; round 1's headline) ------------------------------------------------
    made, ms = time_launches(torch, batch, work_stream, 20, warm=3)
    ms_b, tot_b = reduce_over_ranks(dist if world > 1 else None, dev, ms, float(made))
    burst_value = tot_b * w.ch / (ms_b * 1e-3) / 1e6

    # ---- sustained timed region ---------------------------------------------------------------------------------------------
    warm_steps = max(3, args.warmup)
    for _ in range(warm_steps * per_step):
        batch.launch()
    barrier()
    launches0 = lib.resampleB200KernelLaunches()
    paths0 = path_counts(lib)
    lib.resampleB200ProfileEnable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out_frames = 0
    check_streams = sorted({0, streams - 1})
    snapshots = {}
    total_launches = args.steps * per_step
    with ClockSampler(local) as clocks:
        barrier()
        wall0 = time.perf_counter()
        ev0.record(work_stream)
        for k in range(total_launches):
            if k == total_launches - 1:           # the state the last launch starts from (two host reads per checked stream)
                snapshots = {s: (batch.ctxs[s].contents.outputOffset, batch.ctxs[s].contents.inputIndex) for s in check_streams}
            out_frames += batch.launch()
        ev1.record(work_stream)
        barrier()
        wall1 = time.perf_counter()
    wall_ms = (wall1 - wall0) * 1e3
    ms = ev0.elapsed_time(ev1)
    assert ms > 0.5 * wall_ms or wall_ms < 5.0, f"event time {ms:.3f} ms vs wall {wall_ms:.3f} ms: events miss the work"
    lib.resampleB200ProfileEnable(0)
    kern_ms = C.c_double(0.0)
    kern_launches = lib.resampleB200ProfileCollect(C.byref(kern_ms))
    launches = lib.resampleB200KernelLaunches() - launches0
    paths1 = path_counts(lib)
    kname = kernel_name(paths0, paths1)
    tensor, periodic = kname == "tensor", kname == "periodic"
    last_ring = (batch.at - 1) % RING
    parity = parity_check(torch, batch, snapshots, last_ring) if rank == 0 else None

    ms_max, total_frames = reduce_over_ranks(dist if world > 1 else None, dev, ms, float(out_frames))
    value = total_frames * w.ch / (ms_max * 1e-3) / 1e6
    clock_summary = clocks.summary(wall0, wall1)

    # ---- the same workload with the tensor-core kernel off: strict fp32 accumulation per window ---------------------------
    lib.resampleB200SetTensorPath(0)
    p0 = path_counts(lib)
    made, ms_s = time_launches(torch, batch, work_stream, 12, warm=2)
    strict_kernel = kernel_name(p0, path_counts(lib))
    lib.resampleB200SetTensorPath(1)
    ms_s_max, tot_s = reduce_over_ranks(dist if world > 1 else None, dev, ms_s, float(made))
    strict_value = tot_s * w.ch / (ms_s_max * 1e-3) / 1e6

    # ---- the tensor-core kernel with two signal digits (the round-1 arithmetic: exact to 2^-24 of a tile's peak) -----------------
    two_digit_value = None
    if tensor:
        lib.resampleB200SetTensorDigits(2)
        made, ms_2 = time_launches(torch, batch, work_stream, 20, warm=3)
        lib.resampleB200SetTensorDigits(3)
        ms_2_max, tot_2 = reduce_over_ranks(dist if world > 1 else None, dev, ms_2, float(made))
        two_digit_value = tot_2 * w.ch / (ms_2_max * 1e-3) / 1e6

    # ---- roofline of the convolution kernel ------------------------------------------------------------------------------------
    bpos = w.bytes_per_output_sample
    per_launch_samples = out_frames * w.ch / max(1, kern_launches)
    kern_avg_ms = kern_ms.value / max(1, kern_launches)
    achieved = per_launch_samples * bpos / (kern_avg_ms * 1e-3) / 1e9
    alg_tflops = per_launch_samples * (4 * w.taps + 3) / (kern_avg_ms * 1e-3) / 1e12
    exe_tflops = per_launch_samples * (2 * 416 if periodic else 4 * 384) / (kern_avg_ms * 1e-3) / 1e12
    kernel = ("art_sinc_umma_kernel (tcgen05.mma, 640 threads, 1 CTA/SM)" if tensor else
              "art_sinc_periodic_kernel<CV=2,256>" if periodic else "art_sinc_generic_kernel<interp,float,CV=2>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": kernel,
                "kernel_ms_per_launch": kern_avg_ms, "kernel_share_of_step": kern_ms.value / ms,
                "algorithmic_bytes_per_output_sample": bpos,
                "algorithmic_bytes_per_launch": per_launch_samples * bpos,
                "fp32_tflops_reference_opcount": alg_tflops,
                "timed": "inside the sustained region (clocks as in `clocks`)"}
    if tensor:
        # what the tensor pipe executes: per tile of 128 rows (64 periods x 2 channels), 36 k-steps of 6 MMAs (128 x 160 x 16) --
        # the fixed-point split (6 digit products) and the band's zero blocks (576 executed taps for 380) included
        per_stream_out = out_frames / max(1, total_launches) / streams
        tiles = -(-(-(-per_stream_out // 160)) // 128) * w.ch * streams
        mma_flops = tiles * 36 * 6 * 2.0 * 128 * 160 * 16
        tens = mma_flops / (kern_avg_ms * 1e-3) / 1e12
        tpeak_b = peaks.get("bf16_tflops") or 1622.6
        tpeak_s = peaks.get("bf16_tflops_sustained") or tpeak_b
        roofline.update({
            "tensor_tflops_executed": tens, "tensor_peak_tflops": tpeak_s, "tensor_frac": tens / tpeak_s,
            "tensor_peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (this is a sustained region; fp16 runs at the bf16 rate)",
            "tensor_frac_of_burst_peak": tens / tpeak_b,
            "note": "arithmetic intensity ~200 flop/B puts this path far above the ridge: the kernel is bound by the tensor "
                    "pipe, not by HBM; frac is the HBM-roofline fraction BASELINE.json's metric asks for, tensor_frac the share "
                    "of the measured dense fp16/bf16 tensor throughput the MMAs reach"})
    else:
        roofline.update({
            "fp32_tflops_executed": exe_tflops, "fp32_fma_peak_tflops_at_max_clock": 74.4, "fp32_frac_executed": exe_tflops / 74.4,
            "note": "arithmetic intensity ~200 flop/B puts this path above the FP32 ridge (~11 flop/B): the FP32 FMA "
                    "pipe binds, not HBM; frac is the HBM-roofline fraction BASELINE.json's metric asks for"})
    for cand in ("r02_umma_ncu.json", "r01_umma_ncu.json") if tensor else ("r01_periodic_final_ncu.json",) if periodic else ("r01_generic_v2_ncu.json",):
        prof = ROOT / "profiles" / cand
        if prof.exists():
            try:
                digest = json.loads(prof.read_text())
                captured = float(digest.get("launch_output_samples", 64 * 285344 * 2))      # geometry of the captured launch
                roofline["traffic"] = digest.get("dram_bytes_per_launch") * per_launch_samples / captured
                roofline["traffic_source"] = (f"profiles/{prof.name} (ncu --set full of this kernel on a launch of {captured / 1e6:.1f} M output "
                                              "samples, scaled to this launch by its algorithmic bytes)")
                break
            except Exception:
                pass
    batch.close()

    # ---- end to end through the host-pointer API -----------------------------------------------------------------------------
    e2e = measure_e2e(lib, pkg, torch, dist, world, dev, streams, args.e2e_frames, args) if not args.no_e2e else None

    # ---- every BASELINE config and preset ------------------------------------------------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = measure_configs(lib, pkg, torch, dist, world, rank, dev, work_stream, peak, args)
        configs.append(measure_wide(pkg, torch, dist, world, rank, dev, work_stream, peak, args))

    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm_steps, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"stereo float32 44.1k->48k preset -3 (380x380, interpolated, resampleInit): "
                               f"{streams} independent stereo streams per GPU, a step advances each by {per_step} blocks of {frames} "
                               f"input frames, one batched launch per block",
                   "streams_per_gpu": streams, "frames_per_block": frames, "launches_per_step": per_step,
                   "l2": f"ring of {RING} block sets of {streams * frames * w.ch * 4 / 2**20:.0f} MiB in + "
                         f"{streams * batch.cap * w.ch * 4 / 2**20:.0f} MiB out: every launch touches buffers larger than the 126 MB L2 that "
                         "three other launches have used since",
                   "parallelism": f"streams sharded over {world} GPU(s), no data-path collective",
                   "arithmetic": ("float32 in, float32 out; tensor-core kernel: fixed-point fp16 digit products (3 signal x 3 filter digits, 6 MMAs per 16 taps) "
                                  "with exact fp32 accumulation of the leading term: every sample keeps >= 22 bits of its own magnitude (float accuracy "
                                  "relative to the local signal level); strict_fp32_value is the FFMA form, two_digit_value the round-1 arithmetic" if tensor else
                                  "float32 FMA")},
        "timed_region_s": ms_max * 1e-3, "wall_ms_per_step": wall_ms / args.steps,
        "burst_value": burst_value, "burst": "20 launches on the cold GPU before the sustained region",
        "strict_fp32_value": strict_value, "strict_fp32_kernel": strict_kernel,
        "two_digit_value": two_digit_value,
        "two_digit": "resampleB200SetTensorDigits(2): five MMAs per 16 taps instead of six; exact to 2^-24 of a tile's peak instead of 2^-22 of every sample",
        "clocks": clock_summary, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "parity_check": parity, "configs": configs,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            cores = host_cores()
            s, dt, kind, calls = cpu_run(Workload("metric", streams=cores, frames=16384, **METRIC_WORKLOAD), cores, seconds=8.0)
            line["cpu_baseline"] = {"value": s / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
                                    "sample": f"{cores} stereo streams (one per host thread) x {calls} calls of 16384 input frames "
                                              f"({dt:.1f} s, threads started before the clock), same config"}
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            raise SystemExit(f"parity check failed: {parity}")
    if world > 1:
        dist.destroy_process_group()


def measure_configs(lib, pkg, torch, dist, world, rank, dev, stream, peak, args):
    """BASELINE.json configs 1-5 and presets -1..-4, device-resident, a few launches each (burst clocks), with the HBM-roofline
    fraction and -- on one GPU -- the CPU path's number for the same configuration beside it.  Config 4's 1024 contexts and
    config 3's 64 channels are what the ranks split between them; everything else is replicated per rank (weak)."""
    cores = host_cores()
    n4 = shard_streams(1024, world, rank)[1]
    ch3 = max(1, 64 // world)
    table = [
        ("cfg1 mono preset -1 44.1->48k", Workload("cfg1", 1, 1, 44100, 48000, 64, 1 << 20), 10),
        ("cfg2 stereo preset -3 44.1->48k (metric config)", Workload("cfg2", 2, 3, 44100, 48000, 64, 1 << 18), 10),
        ("cfg2 via resampleFixedRatioInit (art.c:827: 160 filters, no interpolation)", Workload("cfg2f", 2, 3, 44100, 48000, 64, 1 << 18, fixed=True), 6),
        ("cfg2 via resampleFixedRatioInit, tensor-core kernel by opt-in (resampleB200SetTensorPath(3): output no longer bit-identical across call chunkings)",
         Workload("cfg2ft", 2, 3, 44100, 48000, 64, 1 << 18, fixed=True, tensor_mode=3), 8),
        (f"cfg3 {ch3} of 64 ch per GPU, preset -4 96->44.1k, lowpass 20 kHz, two-stage biquad pre-filter (art.c:848-851) folded into the bank",
         Workload("cfg3", ch3, 4, 96000, 44100, 1, 1 << 19, lowpass_hz=20000, biquad="fused"), 6),
        (f"cfg3 with the pre-filter as separate biquad_apply_cascade_interleaved_device calls (the reference's call pattern)",
         Workload("cfg3sep", ch3, 4, 96000, 44100, 1, 1 << 19, lowpass_hz=20000, biquad=True), 4),
        (f"cfg3 without the biquad pre-filter", Workload("cfg3nb", ch3, 4, 96000, 44100, 1, 1 << 19, lowpass_hz=20000), 6),
        (f"cfg3 via resampleFixedRatioInit (art.c:827: 147 filters, no interpolation, automatic lowpass)",
         Workload("cfg3f", ch3, 4, 96000, 44100, 1, 1 << 19, fixed=True), 3),
        (f"cfg3 via resampleFixedRatioInit, tensor-core kernel by opt-in (resampleB200SetTensorPath(3))",
         Workload("cfg3ft", ch3, 4, 96000, 44100, 1, 1 << 19, fixed=True, tensor_mode=3), 6),
        (f"cfg4 {n4} of 1024 stereo contexts per GPU, preset -3 48->44.1k lowpass 20 kHz, 2^15-frame blocks",
         Workload("cfg4s", 2, 3, 48000, 44100, n4, 1 << 15, lowpass_hz=20000), 6),
        (f"cfg4 {n4} of 1024 stereo contexts per GPU, 2^18-frame blocks",
         Workload("cfg4l", 2, 3, 48000, 44100, n4, 1 << 18, lowpass_hz=20000), 4 if n4 > 256 else 8),
        ("cfg5 8 ch preset -2 ASRC +/-100 ppm, 256 blocks x 4096 frames per launch (block API)",
         Workload("cfg5a", 8, 2, 48000, 48000, 1, 4096, asrc_blocks=256, ratio=1.0), 10),
        ("cfg5 8 ch preset -2 ASRC +/-100 ppm, 1024 blocks x 480 frames per launch (block API)",
         Workload("cfg5b", 8, 2, 48000, 48000, 1, 480, asrc_blocks=1024, ratio=1.0), 10),
        ("stereo preset -1 44.1->48k", Workload("p1", 2, 1, 44100, 48000, 64, 1 << 18), 10),
        ("stereo preset -2 44.1->48k", Workload("p2", 2, 2, 44100, 48000, 64, 1 << 18), 10),
        ("stereo preset -3 44.1->48k", Workload("p3", 2, 3, 44100, 48000, 64, 1 << 18), 10),
        ("stereo preset -4 44.1->48k", Workload("p4", 2, 4, 44100, 48000, 64, 1 << 18), 10),
    ]
    out = []
    for name, w, launches in table:
        ring = 2 if w.streams * w.frames * max(1, w.asrc_blocks) * w.ch * 4 > (64 << 20) else 4
        b = DeviceBatch(lib, pkg, torch, dev, w, ring, 777 + rank, stream)
        p0 = path_counts(lib)
        lib.resampleB200SetTensorPath(w.tensor_mode)
        made, ms = time_launches(torch, b, stream, launches, warm=2)
        lib.resampleB200SetTensorPath(1)
        kern = kernel_name(p0, path_counts(lib))
        ms_max, tot = reduce_over_ranks(dist if world > 1 else None, dev, ms, float(made))
        gs = tot * w.ch / (ms_max * 1e-3) / 1e9
        row = {"config": name, "Gsamples_per_s": gs, "ms_per_launch": ms_max / launches, "kernel": kern,
               "hbm_frac": gs * w.bytes_per_output_sample / peak, "filters": int(lib.resampleGetNumFilters(b.ctxs[0])),
               "interpolated": bool(lib.resampleInterpolationUsed(b.ctxs[0]))}
        b.close()
        if world == 1 and rank == 0 and not args.no_cpu:
            wc = Workload(w.name, 64 if w.name.startswith("cfg3") else w.ch, w.preset, w.src, w.dst, cores, w.frames, lowpass_hz=w.lowpass_hz, fixed=w.fixed,
                          biquad=w.biquad, asrc_blocks=w.asrc_blocks, ratio=w.ratio if w.asrc_blocks else None)
            mt = w.name.startswith("cfg3")        # 64 channels: the reference's own RESAMPLE_MULTITHREADED mode, one context
            s, dt, kind, calls = cpu_run(wc, 1 if mt else cores, seconds=1.0, block=4096 if mt else 16384, multithreaded_flag=mt)
            row["cpu_Msamples_per_s"] = s / dt / 1e6
            row["cpu"] = (f"{kind}, RESAMPLE_MULTITHREADED, one 64-channel context" if mt else f"{kind}, {cores} contexts on {cores} threads") + f", {dt:.1f} s"
        out.append(row)
    return out


def measure_wide(pkg, torch, dist, world, rank, dev, stream, peak, args):
    """One row for the PATH_WIDTH=64 library (libresampler_b200_64.so: double samples): the metric's
    configuration, device-resident, and -- on one GPU -- the reference's own PATH_WIDTH=64 build on the host beside it."""
    lib64 = pkg.load64()
    ch, (filters, taps), ratio = 2, PRESETS[3], 48000 / 44100
    n, frames = 64, 1 << 16
    lib64.resampleB200SetDevice(dev.index or 0)
    ctxs = [lib64.resampleInit(ch, taps, filters, 0.0, FLAGS) for _ in range(n)]
    assert all(ctxs), "resampleInit (PATH_WIDTH=64) failed"
    for c in ctxs:
        lib64.resampleAdvancePosition(c, taps / 2)
    x = torch.rand((n, frames, ch), device=dev, dtype=torch.float64) - 0.5
    cap = int(frames * ratio) + taps + 16
    y = torch.empty((n, cap, ch), device=dev, dtype=torch.float64)
    ca = (C.POINTER(pkg.Resample64) * n)(*ctxs)
    ia = (C.c_void_p * n)(*[x[i].data_ptr() for i in range(n)]); oa = (C.c_void_p * n)(*[y[i].data_ptr() for i in range(n)])
    ni, no = (C.c_int * n)(*([frames] * n)), (C.c_int * n)(*([cap] * n))
    ra = (C.c_double * n)(*([ratio] * n)); res = (pkg.ResampleResult * n)()
    sp = C.c_void_p(stream.cuda_stream)

    def step():
        lib64.resampleBatchProcessInterleavedDevice(ca, n, ia, ni, oa, no, ra, res, sp)
        return int(np.frombuffer(res, dtype=np.uint32)[1::2].sum(dtype=np.int64))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); made = 0
    launches = 5
    for _ in range(launches):
        made += step()
    e1.record(stream); torch.cuda.synchronize()
    ms_max, tot = reduce_over_ranks(dist if world > 1 else None, dev, e0.elapsed_time(e1), float(made))
    gs = tot * ch / (ms_max * 1e-3) / 1e9
    row = {"config": "PATH_WIDTH=64 library (double samples): stereo preset -3 44.1->48k, 64 streams x 2^16 frames", "Gsamples_per_s": gs,
           "ms_per_launch": ms_max / launches, "kernel": "generic (double multiply-adds)", "hbm_frac": gs * 8.0 * (1.0 + 1.0 / ratio) / peak,
           "filters": filters, "interpolated": True}
    for c in ctxs:
        lib64.resampleFree(c)
    del x, y
    torch.cuda.empty_cache()
    if world == 1 and rank == 0 and not args.no_cpu:
        import artlibs as A
        if A.reference64() is not None:
            cores = host_cores()
            block = 16384
            xin = np.random.default_rng(5).uniform(-0.5, 0.5, (block, ch))
            streams = [A.reference_stream64(ch, taps, filters, 0.0, FLAGS) for _ in range(cores)]
            for s in streams:
                s.advance(taps / 2)
            made_by = [0] * cores
            calls = 40
            go = threading.Barrier(cores + 1)

            def work(t):
                s = streams[t]
                out = np.empty((int(block * ratio) + taps + 16, ch))
                fn, xp, op = s.lib.resampleProcessInterleaved, xin.ctypes.data_as(A.f64p), out.ctypes.data_as(A.f64p)
                fn(s.ctx, xp, block, op, out.shape[0], ratio)
                go.wait()
                tot_ = 0
                for _ in range(calls):
                    tot_ += fn(s.ctx, xp, block, op, out.shape[0], ratio).output_generated
                made_by[t] = tot_
                go.wait()
            pool = [threading.Thread(target=work, args=(t,), daemon=True) for t in range(cores)]
            for th in pool:
                th.start()
            go.wait(); t0 = time.perf_counter(); go.wait(); dt = time.perf_counter() - t0
            for th in pool:
                th.join()
            row["cpu_Msamples_per_s"] = sum(made_by) * ch / dt / 1e6
            row["cpu"] = f"reference built with -DPATH_WIDTH=64, {cores} contexts on {cores} threads, {dt:.1f} s"
    return row


def measure_e2e(lib, pkg, torch, dist, world, dev, streams, frames, args):
    """Same workload with pinned HOST buffers, host->device and device->host copies inside the timed region.

    value          through the reference-facing call, resampleProcessInterleaved (include/resampler.h), one
                   context per call, a few host threads keeping several contexts in flight (they sleep on a
                   blocking-sync event while their transfers run: no core is spent polling);
    batched_value  through the host-pointer batch extension (resampleBatchProcessInterleaved), one call per step.
    """
    w = Workload("metric", streams=min(streams, args.e2e_streams), frames=frames, **METRIC_WORKLOAD)
    n, ch = w.streams, w.ch
    cap = int(frames * w.ratio) + w.taps + 16
    hx = torch.empty((n, frames, ch), dtype=torch.float32).uniform_(-0.5, 0.5).pin_memory()
    hy = torch.empty((n, cap, ch), dtype=torch.float32).pin_memory()
    f32p = C.POINTER(C.c_float)
    xp = [C.cast(hx[i].data_ptr(), f32p) for i in range(n)]
    yp = [C.cast(hy[i].data_ptr(), f32p) for i in range(n)]
    local = dev.index
    steps = args.e2e_steps                        # ~1 s at the PCIe floor of one GPU

    def fresh():
        ctxs = [lib.resampleInit(ch, w.taps, w.filters, 0.0, FLAGS) for _ in range(n)]
        for c in ctxs:
            lib.resampleAdvancePosition(c, w.taps / 2)
        return ctxs

    def reduce(made, dt):
        t, tot = reduce_over_ranks(dist if world > 1 else None, dev, dt, float(made))
        return tot * ch / t / 1e6

    # -- reference-facing API: long-lived host threads, each owning every T-th stream --------------------------------
    ctxs = fresh()
    T = max(1, min(args.e2e_threads or min(16, max(4, host_cores() // max(world, 1))), n))
    start, done = threading.Barrier(T + 1), threading.Barrier(T + 1)
    made_by = [0] * T
    rounds = {"n": 0}

    def worker(t):
        lib.resampleB200SetDevice(local)
        fn = lib.resampleProcessInterleaved
        mine = [(ctxs[i], xp[i], yp[i]) for i in range(t, n, T)]
        while True:
            start.wait()
            if rounds["n"] < 0:
                return
            tot = 0
            for _ in range(rounds["n"]):
                for c, xi, yi in mine:
                    tot += fn(c, xi, frames, yi, cap, w.ratio).output_generated
            made_by[t] = tot
            done.wait()

    threads = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(T)]
    for th in threads:
        th.start()

    def run_rounds(k):
        rounds["n"] = k
        start.wait()
        done.wait()
        return sum(made_by)

    run_rounds(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    made = run_rounds(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rounds["n"] = -1
    start.wait()
    for th in threads:
        th.join()
    value = reduce(made, dt)
    per_step_out = made / steps
    for c in ctxs:
        lib.resampleFree(c)

    # -- host-pointer batch extension, one call per step: ONE submitting thread per rank ------------------------------
    ctxs = fresh()
    ctx_t = C.POINTER(pkg.Resample)
    ctx_arr = (ctx_t * n)(*ctxs)
    in_arr, out_arr = (f32p * n)(*xp), (f32p * n)(*yp)
    nin_arr, nout_arr = (C.c_int * n)(*([frames] * n)), (C.c_int * n)(*([cap] * n))
    ratio_arr = (C.c_double * n)(*([w.ratio] * n))
    res_arr = (pkg.ResampleResult * n)()

    def batch():
        lib.resampleBatchProcessInterleaved(ctx_arr, n, in_arr, nin_arr, out_arr, nout_arr, ratio_arr, res_arr)
        return int(np.frombuffer(res_arr, dtype=np.uint32)[1::2].sum(dtype=np.int64))

    for _ in range(2):
        batch()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    made = 0
    for _ in range(steps):
        made += batch()
    torch.cuda.synchronize()
    dt_b = time.perf_counter() - t0
    batched = reduce(made, dt_b)
    for c in ctxs:
        lib.resampleFree(c)

    return {"value": value, "unit": "Msamples/s",
            "h2d_bytes_per_step": int(n * frames * ch * 4), "d2h_bytes_per_step": int(per_step_out * ch * 4),
            "api": "resampleProcessInterleaved (host pointers, pinned), "
                   f"{n} streams x {frames} frames per step, {T} host threads per rank (the library waits spinning while the process has a CPU per waiting thread, ART_B200_WAIT), {steps} steps in {dt:.2f} s",
            "batched_value": batched,
            "batched_api": "resampleBatchProcessInterleaved (host pointers, pinned), one call per step, one host thread per rank"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=74,
                    help="independent stereo streams per GPU (74 x 28 tiles of 64 periods = 2072 = 14 per SM: no partial last wave)")
    ap.add_argument("--frames", type=int, default=1 << 20,
                    help="input frames per stream per block (= per launch): 23.8 s of audio; 74 streams x 112 tiles = 56 tiles per SM")
    ap.add_argument("--launches-per-step", type=int, default=56, help="blocks every stream advances by in one step (~0.9 ms each)")
    ap.add_argument("--e2e-streams", type=int, default=64)
    ap.add_argument("--e2e-threads", type=int, default=0, help="host threads per rank in the end-to-end leg; 0 = the CPUs this rank may use, between 4 and 16")
    ap.add_argument("--e2e-frames", type=int, default=1 << 18, help="input frames per host call in the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
