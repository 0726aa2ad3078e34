#!/usr/bin/env python3
"""Benchmark of the B200 MDEC/BS + ADPCM encode core over the BASELINE.json config matrix.

Headline (BASELINE.json configs[1], `strv`): a "step" encodes one batch of 4096 synthetic
320x240 NV21 frames per GPU to BS v2 bitstreams with a 20160-byte budget per frame (psxavenc
`-t strv` defaults: 15 fps at 2x CD speed = 10 sectors of 2016 bytes per frame,
filefmt.c:540-552; mdec.c:772-774) through the C ABI of libpsxav_b200.so. ONE JSON line is
printed by rank 0:

  value        frames/s, whole job, inputs resident in HBM, CUDA events on the launching stream,
               max over ranks, exactly --steps steps
  sustained    the same loop kept running for >= 1 s (clocks sampled throughout)
  e2e          the same metric through the host-buffer C-ABI entry points with pinned HOST
               buffers (host->device and device->host copies inside the timed region)
  roofline     whole step and per kernel: algorithmic bytes / CUDA-event duration against the
               measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline / cpu_baseline_1thread / --impl reference
               the UNMODIFIED reference C (oracle/_ref/libpsxav_ref.so; the oracle port when that
               cannot be loaded) on the host's cores, bounded sample
  configs      the other BASELINE configs, each measured and parity-checked in the same run:
               strv easy/hard content, sbs, strcd (video + XA concurrently), vagi (B=1, x1024),
               spu (the sine, one block per call through the drop-in symbol)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bs_v2_320x240_frames_per_sec"
UNIT = "frames/s"
FRAMES_PER_STEP = 4096            # per GPU (weak scaling)


class Workload:
    """One BS video shape of BASELINE.json."""

    def __init__(self, name, width, height, codec, max_size, frames, noise=3, smooth=False):
        self.name, self.width, self.height, self.codec, self.max_size = name, width, height, codec, max_size
        self.frames, self.noise, self.smooth = frames, noise, smooth
        self.frame_bytes = width * height * 3 // 2
        # NV21 read + the whole bitstream buffer written
        self.algo_bytes = self.frame_bytes + max_size

    def describe(self):
        content = "smooth sinusoidal content" if self.smooth else "noise_bits=%d" % self.noise
        return "%s: %dx%d BS %s, frame_max_size %d B, %d synthetic NV21 frames per step per GPU, %s" % (
            self.name, self.width, self.height, ["v2", "v3", "v3dc"][self.codec], self.max_size, self.frames, content)

    def host_frames(self, first, distinct=256):
        """`distinct` integer-generator frames tiled to the batch; the
        copies are made unique on the device by the caller."""
        from psxavenc_b200 import synth
        count = self.frames
        if self.smooth:
            base = np.stack([synth.gen_smooth_frame(first + i, self.width, self.height, amplitude=20 + 4 * (i % 24),
                                                    fx=0.006 + 0.001 * (i % 13), fy=0.009 + 0.0007 * (i % 11))
                             for i in range(min(distinct // 4, count))])
        else:
            base = synth.gen_frames(first, min(distinct, count), self.width, self.height, self.noise)
        reps = (count + len(base) - 1) // len(base)
        return np.tile(base, (reps, 1))[:count]


def headline_config(fdct_name):
    """`config` of the JSON line — the same dict in both arms, so that the driver can tell they ran
    the same thing; what only concerns one arm is in `config_detail`."""
    return {"workload": WORKLOAD, "fdct": fdct_name,
            "l2": "GPU arm: inputs larger than L2 (%.0f MB of frames per step per GPU, every frame distinct); "
                  "reference arm: the same frames in host memory" % (FRAMES_PER_STEP * 115200 / 1e6)}


STRV = Workload("strv", 320, 240, 0, 20160, FRAMES_PER_STEP, noise=3)
WORKLOAD = "strv: 320x240 BS v2, frame_max_size 20160 B, %d synthetic NV21 frames per step per GPU, noise_bits=3 (quant scale 2)" % FRAMES_PER_STEP


def fdct_from_name(name):
    return 1 if name == "sse2" else 0


# ---------------------------------------------------------------------------------------
# CPU side: the reference (or the oracle port) on the host's cores
# ---------------------------------------------------------------------------------------

def cpu_backend():
    import oracle
    try:
        return oracle.Reference()
    except Exception:
        return oracle.Restated()


class CpuEncoder:
    """Runs the CPU implementation of the path over slices of a frame batch on `cores` threads
    (ctypes releases the GIL; every thread owns its encoder handle)."""

    def __init__(self, wl, fdct, cores=None):
        self.wl, self.fdct = wl, fdct
        self.cores = cores or os.cpu_count() or 1
        self.backend = cpu_backend()
        self.kind = self.backend.kind

    def encode(self, frames, sizes=None):
        wl, n = self.wl, len(frames)
        chunks = [c for c in np.array_split(np.arange(n), min(self.cores, n)) if len(c)]
        results = [None] * len(chunks)

        def work(i):
            lo, hi = chunks[i][0], chunks[i][-1] + 1
            budget = wl.max_size if sizes is None else sizes[lo:hi]
            results[i] = self.backend.bs_encode_batch(wl.codec, wl.width, wl.height, frames[lo:hi], budget, self.fdct,
                                                      stride=wl.max_size)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(chunks))]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        return dt, np.concatenate([r[0] for r in results]), np.concatenate([r[1] for r in results])

    def calibrate(self, frames, target_seconds):
        """Frames per bounded sample so that one sample costs about target_seconds."""
        probe = frames[:max(self.cores, 8)]
        dt, _, _ = self.encode(probe)
        n = int(len(probe) / dt * target_seconds)
        return max(self.cores, min(len(frames), n // self.cores * self.cores))


def cpu_baseline(wl, frames, fdct, target_seconds=1.5, cores=None, max_frames=None):
    """`cores` host threads (default all) on (a prefix of) the step batch, repeated until about
    target_seconds of wall time. Returns the baseline object and the outputs of the last pass."""
    cpu = CpuEncoder(wl, fdct, cores)
    if max_frames is None:
        max_frames = cpu.calibrate(frames, target_seconds) if cores else len(frames)
    sample = frames[:max_frames]
    total, reps, out, res = 0.0, 0, None, None
    while total < target_seconds and reps < 64:
        dt, out, res = cpu.encode(sample)
        total += dt
        reps += 1
    return {
        "value": reps * len(sample) / total, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
        "sample": "%d frames of the step batch x %d passes, %d threads, %.1f s wall" % (len(sample), reps, cpu.cores, total),
    }, out, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fdct = fdct_from_name(args.fdct)
    cpu = CpuEncoder(STRV, fdct)
    frames = STRV.host_frames(0)
    n = cpu.calibrate(frames, 2.5)
    for _ in range(args.warmup):
        cpu.encode(frames[:n])
    total = 0.0
    for _ in range(args.steps):
        dt, _, _ = cpu.encode(frames[:n])
        total += dt
    value = n * args.steps / total
    sample = "%d of the %d frames of a step per step, %d threads" % (n, FRAMES_PER_STEP, cpu.cores)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": headline_config(args.fdct), "config_detail": {"sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML (in-process, every few
    milliseconds) while a timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.004):
        self.samples, self.bits, self.max_mhz, self.power = [], 0, None, []
        self.stop_flag = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, args=(period,), daemon=True)
            self.thread.start()
        except Exception as e:   # pragma: no cover
            self.error = str(e)

    def _run(self, period):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(period)

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % getattr(self, "error", "?")]}
        self.stop_flag.set()
        self.thread.join()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None,
                "reasons": reasons}


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the kernels from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class Ctx:
    """Per-process GPU context shared by the legs."""

    def __init__(self, args):
        import torch
        import psxavenc_b200 as pb
        self.torch, self.pb, self.args = torch, pb, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available() or pb.device_count() == 0:
            raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        self.host_group = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
            # a second, CPU-side group: ranks that wait on it leave their GPU idle (an NCCL barrier
            # parks a spinning kernel on it), which the single-process multi-device leg needs
            self.host_group = dist.new_group(backend="gloo")
        self.stream = torch.cuda.current_stream()
        self.fdct = fdct_from_name(args.fdct)
        self.fdct_name = args.fdct
        self.sm_count = torch.cuda.get_device_properties(self.local).multi_processor_count
        self.peak, self.peak_src = load_peak()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        self.torch.cuda.synchronize()
        if self.host_group is not None:
            self.dist.barrier(group=self.host_group)

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def all_ranks_ok(self, ok):
        t = self.torch.tensor([0 if ok else 1], dtype=self.torch.int32, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item()) == 0

    def timed(self, fn, steps, min_seconds=0.0):
        """fn() `steps` times (then on until min_seconds), CUDA events on the launching stream,
        barrier + synchronize on both sides. -> (ms, passes)."""
        torch = self.torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        a.record(self.stream)
        passes = 0
        t0 = time.perf_counter()
        while passes < steps or (min_seconds and time.perf_counter() - t0 < min_seconds and passes < 100000):
            fn()
            passes += 1
            if min_seconds and passes >= steps and passes % 8 == 0:
                self.stream.synchronize()       # keep the launch queue bounded while watching the clock
        b.record(self.stream)
        self.barrier()
        return a.elapsed_time(b), passes


class VideoLeg:
    """Device-resident frames of one workload + the encoder and output buffers."""

    def __init__(self, ctx, wl, max_batch=None, distinct=256):
        torch, pb = ctx.torch, ctx.pb
        self.ctx, self.wl = ctx, wl
        n = wl.frames
        host = wl.host_frames(ctx.rank * n, distinct)
        self.d_frames = torch.from_numpy(host).to(ctx.dev)
        if not wl.smooth and n > distinct:
            # flip the lowest luma bit of the tiled copies so that no two frames are identical
            gen = torch.Generator(device=ctx.dev)
            gen.manual_seed(1234 + ctx.rank + 7 * wl.noise)
            salt = torch.randint(0, 2, (n, wl.width * wl.height), dtype=torch.uint8, device=ctx.dev, generator=gen)
            salt[:distinct] = 0
            self.d_frames[:, :wl.width * wl.height] ^= salt
            del salt
        self.d_sizes = torch.full((n,), wl.max_size, dtype=torch.int32, device=ctx.dev)
        self.d_out = torch.zeros((n, wl.max_size), dtype=torch.uint8, device=ctx.dev)
        self.d_res = torch.zeros((n, 4), dtype=torch.int32, device=ctx.dev)
        self.enc = pb.BsEncoder(wl.codec, wl.width, wl.height, ctx.fdct, max_batch=max_batch or n)

    def step(self):
        wl = self.wl
        self.enc.encode_device(wl.frames, self.d_frames, self.d_sizes, wl.max_size, self.d_out, wl.max_size, self.d_res,
                               self.ctx.stream.cuda_stream)

    def close(self):
        self.enc.close()


def check_against_cpu(leg, cpu_out, cpu_res, what):
    k = len(cpu_res)
    res = leg.d_res[:k].cpu().numpy()
    if not (np.array_equal(cpu_res, res) and np.array_equal(cpu_out, leg.d_out[:k].cpu().numpy())):
        raise SystemExit("bench.py: GPU output differs from the CPU reference's output (%s) — number withheld" % what)
    return k


def bench_headline(ctx, line):
    """strv, typical content: value, sustained, roofline, e2e, cpu baselines."""
    torch, pb, args = ctx.torch, ctx.pb, ctx.args
    wl = STRV
    n = wl.frames
    leg = VideoLeg(ctx, wl, max_batch=args.chunk)
    enc = leg.enc
    gathered = torch.zeros((ctx.world * n, 4), dtype=torch.int32, device=ctx.dev) if ctx.world > 1 else None
    side = torch.cuda.Stream() if ctx.world > 1 else None
    gather_done = torch.cuda.Event() if ctx.world > 1 else None

    def step():
        leg.step()
        if ctx.world > 1:
            # the path's only exchange: per-frame {bytes_used, blocks_used, q, hwords} to the muxing
            # rank — on a side stream, so that the latency-bound collective of step k hides behind
            # the kernels of step k + 1
            gather_done.record(ctx.stream)
            side.wait_event(gather_done)
            with torch.cuda.stream(side):
                ctx.dist.all_gather_into_tensor(gathered, leg.d_res)

    def drain():
        if side is not None:
            ctx.stream.wait_stream(side)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    drain()
    ctx.barrier()

    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    launches0 = pb.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    a.record(ctx.stream)
    for _ in range(args.steps):
        step()
    drain()
    b.record(ctx.stream)
    ctx.barrier()
    elapsed_ms = a.elapsed_time(b)
    launches = pb.launch_count() - launches0

    # the same loop for at least a second: sustained clocks
    sus_ms, sus_passes = ctx.timed(lambda: (step(), drain()), args.steps, min_seconds=1.0)
    clocks = sampler.stop() if sampler else None

    # per-kernel durations: a few more steps with every launch bracketed by CUDA events on its
    # stream (the library then issues the launches of a step back to back on one stream)
    enc.timing(True)
    enc.read_timing()
    for _ in range(5):
        leg.step()
    torch.cuda.synchronize()
    dct_ms, pack_ms, pairs = enc.read_timing()
    enc.timing(False)

    # ---- parity spot check against the CPU oracle on this run's own bytes (every rank) --------
    import oracle
    res = leg.d_res.cpu().numpy()
    sel = np.array([0, 1, 255, 256, 257, min(1000, n - 2), n - 1])
    idx = torch.from_numpy(sel).to(ctx.dev)
    exp_out, exp_res = oracle.Restated().bs_encode_batch(wl.codec, wl.width, wl.height, leg.d_frames[idx].cpu().numpy(),
                                                         wl.max_size, ctx.fdct)
    ok = bool(np.array_equal(leg.d_out[idx].cpu().numpy(), exp_out) and np.array_equal(res[sel], exp_res))
    if ctx.world > 1:
        ok = ok and bool(torch.equal(gathered[ctx.rank * n:(ctx.rank + 1) * n], leg.d_res))
    if not ctx.all_ranks_ok(ok):
        raise SystemExit("bench.py: GPU output differs from the CPU oracle on some rank — number withheld")

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ------------------
    h_frames = torch.empty((n, wl.frame_bytes), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(leg.d_frames)
    h_sizes = torch.full((n,), wl.max_size, dtype=torch.int32).pin_memory()
    h_out = torch.empty((n, wl.max_size), dtype=torch.uint8, pin_memory=True)
    h_res = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
    host_enc = pb.BsEncoder(wl.codec, wl.width, wl.height, ctx.fdct, max_batch=256)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        host_enc.encode_host_into(n, h_frames, h_sizes, h_out, wl.max_size, h_res)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        failed = host_enc.encode_host_into(n, h_frames, h_sizes, h_out, wl.max_size, h_res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ok = failed == 0 and np.array_equal(h_res.numpy(), res) and torch.equal(h_out, leg.d_out.cpu())
    if not ctx.all_ranks_ok(ok):
        raise SystemExit("bench.py: e2e output differs from the device-resident output")
    d2h_bytes = int(((res[:, 0] + 63) // 64 * 64).max()) * n + 16 * n      # what the compacted copy moves
    host_enc.close()

    # the same host->device link measured bare (pinned cudaMemcpy of the step's frames): the e2e
    # path is bound by it, so its share of this figure is what the pipeline can be judged by
    d_probe = torch.empty_like(leg.d_frames)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_probe.copy_(h_frames, non_blocking=True)
    p0.record(ctx.stream)
    for _ in range(3):
        d_probe.copy_(h_frames, non_blocking=True)
    p1.record(ctx.stream)
    torch.cuda.synchronize()
    link_gbs = 3 * n * wl.frame_bytes / (p0.elapsed_time(p1) / 1000.0) / 1e9
    del d_probe

    elapsed_ms, e2e_ms, dct_ms, pack_ms, sus_ms = ctx.max_over_ranks([elapsed_ms, e2e_s * 1000.0, dct_ms, pack_ms, sus_ms])

    # ---- e2e through ONE process driving all GPUs (rank 0; the others wait) -------------------
    single = None
    if ctx.world > 1 and pb.device_count() >= ctx.world:
        ctx.host_barrier()
        if ctx.rank == 0:
            try:
                single = e2e_single_process(ctx, wl, h_frames.numpy(), res, h_out.numpy())
            except Exception as e:      # keep the per-rank number if the single-process leg cannot run
                single = {"error": str(e)}
        ctx.host_barrier()

    if ctx.rank == 0:
        world = ctx.world
        value = world * n * args.steps / (elapsed_ms / 1000.0)
        frames_per_launch = min(args.chunk, n)
        traffic = load_traffic()
        step_ms = elapsed_ms / args.steps
        per_kernel = {}
        for name, ms, algo in (("bs_dct_kernel", dct_ms / max(pairs, 1), wl.frame_bytes),
                               ("bs_pack_kernel", pack_ms / max(pairs, 1), wl.max_size)):
            gbs = algo * frames_per_launch / (ms / 1000.0) / 1e9 if ms > 0 else 0.0
            per_kernel[name] = {"algorithmic_bytes_per_launch": algo * frames_per_launch, "launch_ms": ms, "gbs": gbs,
                                "frac": gbs / ctx.peak, "traffic": traffic.get(name),
                                "share_of_kernel_time": ms * pairs / (dct_ms + pack_ms) if dct_ms + pack_ms else None}
        if ctx.fdct_name == "sse2" and clocks and clocks.get("sm_mhz"):
            # What bounds the FDCT kernel is its own instruction mix (straight-line code: static = executed): 864 ALU-pipe
            # instructions at 2 cycles and 745 IMAD at 2 + 64 IMAD.HI at 4 cycles per warp and sub-partition, the rates measured
            # by tools/ubench/pipes.cu (profiles/r2_microopt_ab.txt); one warp = one group of 32 blocks.
            nmb = (wl.width // 16) * (wl.height // 16)
            warps = frames_per_launch * ((2 * nmb + 31) // 32 + (4 * nmb + 31) // 32)
            cycles = max(864 * 2, 745 * 2 + 64 * 4)
            bound_ms = warps / (4.0 * ctx.sm_count) * cycles / (clocks["sm_mhz"] * 1e3)
            per_kernel["bs_dct_kernel"]["pipe_bound"] = {
                "alu_cycles_per_warp": 864 * 2, "fma_cycles_per_warp": 745 * 2 + 64 * 4, "bound_ms": bound_ms,
                "frac": bound_ms / per_kernel["bs_dct_kernel"]["launch_ms"] if per_kernel["bs_dct_kernel"]["launch_ms"] else None,
                "source": "static SASS mix of bs_dct_kernel<sse2> x measured pipe issue rates (DESIGN.md 4.1)"}
        whole = wl.algo_bytes * n / (step_ms / 1000.0) / 1e9
        e2e_value = world * n * e2e_steps / (e2e_ms / 1000.0)
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * wl.frame_bytes,
               "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
               "timer": "host wall clock around the synchronous C-ABI call, max over ranks",
               "h2d_gbs_per_gpu": n * wl.frame_bytes * e2e_steps / (e2e_ms / 1000.0) / 1e9,
               "h2d_link_gbs_rank0": link_gbs,
               "api": "psxb200_bs_encode_host, one process per GPU (pinned host buffers, compacted copy-back)",
               "per_rank_processes": {"value": e2e_value}}
        if single and "value" in single:
            e2e["single_process"] = single
            if single["value"] > e2e_value:
                e2e.update(value=single["value"], api=single["api"], h2d_gbs_per_gpu=single["h2d_gbs_per_gpu"])
        elif single:
            e2e["single_process"] = single
        line.update({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": headline_config(args.fdct),
            "config_detail": {"frames_per_launch": frames_per_launch,
                              "launch_pipeline": "launches of a step alternate between two forked streams" if n > frames_per_launch
                              and os.environ.get("PSXB200_DEVICE_PIPELINE", "1") != "0" else "one stream",
                              "quant_scale_mean": float(res[:, 2].mean()), "parity_spot_check": True,
                              "collective": "all_gather of per-frame results on a side stream" if world > 1 else "none"},
            "sustained": {"value": world * n * sus_passes / (sus_ms / 1000.0), "unit": UNIT, "passes": sus_passes,
                          "seconds": sus_ms / 1000.0},
            "roofline": {"bound": "hbm", "kernel": "bs_dct_kernel + bs_pack_kernel (whole step)", "achieved": whole,
                         "peak": ctx.peak, "unit": "GB/s", "frac": whole / ctx.peak,
                         "traffic": (traffic.get("bs_dct_kernel") or 0) + (traffic.get("bs_pack_kernel") or 0) or None,
                         "peak_source": ctx.peak_src, "algorithmic_bytes_per_launch": wl.algo_bytes * frames_per_launch,
                         "algorithmic_bytes_per_frame": wl.algo_bytes, "step_ms": step_ms,
                         "per_kernel": per_kernel, "kernel_ms_per_step_serial": (dct_ms + pack_ms) / 5, "step_ms_total": elapsed_ms,
                         "note": "integer-pipe bound kernels (ncu: profiles/r2b_*; FDCT: per_kernel.bs_dct_kernel.pipe_bound): the HBM "
                                 "fraction is reported, not the limiter"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        })
        if world == 1 and not args.no_cpu:
            frames_host = leg.d_frames.cpu().numpy()
            base, cpu_out, cpu_res = cpu_baseline(wl, frames_host, ctx.fdct)
            base["parity_frames_checked"] = check_against_cpu(leg, cpu_out, cpu_res, "strv")
            line["cpu_baseline"] = base
            one, _, _ = cpu_baseline(wl, frames_host, ctx.fdct, target_seconds=1.5, cores=1)
            line["cpu_baseline_1thread"] = one
    leg.close()


def e2e_single_process(ctx, wl, h_frames_rank0, res_rank0, out_rank0):
    """All GPUs of the box from ONE process through psxb200_bs_multi_encode_host (the entry the
    reference's single-process C host would call). Rank 0 only; frames of rank 0 repeated per device."""
    torch, pb = ctx.torch, ctx.pb
    g, n = ctx.world, wl.frames
    total = g * n
    h_frames = torch.empty((total, wl.frame_bytes), dtype=torch.uint8, pin_memory=True)
    for d in range(g):
        h_frames[d * n:(d + 1) * n].copy_(torch.from_numpy(h_frames_rank0))
    h_sizes = torch.full((total,), wl.max_size, dtype=torch.int32).pin_memory()
    h_out = torch.empty((total, wl.max_size), dtype=torch.uint8, pin_memory=True)
    h_res = torch.empty((total, 4), dtype=torch.int32, pin_memory=True)
    multi = pb.BsMultiEncoder(wl.codec, wl.width, wl.height, ctx.fdct, max_batch=256, n_devices=g)
    for _ in range(2):
        multi.encode_host_into(total, h_frames, h_sizes, h_out, wl.max_size, h_res)
    steps = 5
    t0 = time.perf_counter()
    for _ in range(steps):
        failed = multi.encode_host_into(total, h_frames, h_sizes, h_out, wl.max_size, h_res)
    dt = time.perf_counter() - t0
    multi.close()
    ok = failed == 0
    for d in (0, g - 1):
        ok = ok and np.array_equal(h_res.numpy()[d * n:(d + 1) * n], res_rank0) and np.array_equal(h_out.numpy()[d * n:(d + 1) * n], out_rank0)
    if not ok:
        raise SystemExit("bench.py: multi-device e2e output differs from the device-resident output")
    return {"value": total * steps / dt, "unit": UNIT, "devices": g, "steps": steps,
            "h2d_gbs_per_gpu": n * wl.frame_bytes * steps / dt / 1e9,
            "api": "psxb200_bs_multi_encode_host, one process driving %d GPUs (worker thread per device)" % g}


def bench_content(ctx, noise, label):
    """strv with easy (noise 0 -> q = 1) / hard (noise 6 -> q = 8) content."""
    wl = Workload("strv", 320, 240, 0, 20160, FRAMES_PER_STEP, noise=noise)
    leg = VideoLeg(ctx, wl)
    for _ in range(3):
        leg.step()
    steps = max(5, min(ctx.args.steps, 20))
    ms, passes = ctx.timed(leg.step, steps)
    (ms,) = ctx.max_over_ranks([ms])
    res = leg.d_res.cpu().numpy()
    out = {"content": label, "noise_bits": noise, "value": ctx.world * wl.frames * passes / (ms / 1000.0), "unit": UNIT,
           "ms_per_step": ms / passes, "quant_scale_mean": float(res[:, 2].mean())}
    if ctx.rank == 0 and ctx.world == 1 and not ctx.args.no_cpu:
        base, cpu_out, cpu_res = cpu_baseline(wl, leg.d_frames[:1024].cpu().numpy(), ctx.fdct, target_seconds=1.0)
        base["parity_frames_checked"] = check_against_cpu(leg, cpu_out, cpu_res, "strv " + label)
        out["cpu_baseline"] = base
    leg.close()
    return out


def bench_sbs(ctx):
    """BASELINE config `sbs`: 640x480 BS v3, 8192-byte frames; the 10 000-frame batch, sharded in
    contiguous frame ranges over the GPUs (strong scaling: the batch is fixed)."""
    from psxavenc_b200 import sharding
    total = 10000
    first, last = sharding.frame_range(total, ctx.rank, ctx.world)
    wl = Workload("sbs", 640, 480, 1, 8192, last - first, smooth=True)
    leg = VideoLeg(ctx, wl, distinct=384)
    for _ in range(2):
        leg.step()
    steps = 3
    ms, passes = ctx.timed(leg.step, steps)
    (ms,) = ctx.max_over_ranks([ms])
    res = leg.d_res.cpu().numpy()
    ok = bool((res[:, 2] < 64).all())
    out = {"workload": "sbs: 640x480 BS v3, frame_max_size 8192 B, one 10000-frame batch in contiguous ranges over %d GPU(s), "
                       "smooth sinusoidal content" % ctx.world,
           "value": total * passes / (ms / 1000.0), "unit": UNIT, "ms_per_batch": ms / passes, "scaling": "strong",
           "frames_per_gpu": last - first, "quant_scale_mean": float(res[:, 2].mean()),
           "roofline_whole_step_frac": wl.algo_bytes * (last - first) / (ms / passes / 1000.0) / 1e9 / ctx.peak}
    # parity on every rank: a sample of its own frames against the unmodified reference
    backend = cpu_backend()
    sel = np.unique(np.linspace(0, wl.frames - 1, 12).astype(np.int64))
    idx = ctx.torch.from_numpy(sel).to(ctx.dev)
    t0 = time.perf_counter()
    exp_out, exp_res = backend.bs_encode_batch(wl.codec, wl.width, wl.height, leg.d_frames[idx].cpu().numpy(), wl.max_size, ctx.fdct)
    cpu_dt = time.perf_counter() - t0
    ok = ok and np.array_equal(res[sel][:, :3], exp_res[:, :3]) and np.array_equal(leg.d_out[idx].cpu().numpy(), exp_out)
    if not ctx.all_ranks_ok(ok):
        raise SystemExit("bench.py: sbs output differs from the CPU reference")
    out["parity_frames_checked_per_rank"] = len(sel)
    if ctx.rank == 0 and ctx.world == 1 and not ctx.args.no_cpu:
        base, _, _ = cpu_baseline(wl, leg.d_frames[:256].cpu().numpy(), ctx.fdct, target_seconds=1.5)
        out["cpu_baseline"] = base
        out["cpu_baseline_1thread"] = {"value": len(sel) / cpu_dt, "unit": UNIT, "cores": 1, "kind": backend.kind,
                                       "sample": "%d frames, one thread" % len(sel)}
    leg.close()
    return out


def bench_strcd(ctx):
    """BASELINE config `strcd`: 320x240 v2 video with budgets 16128,18144,18144,... (8.75 sectors per
    frame) + 37800 Hz 4-bit stereo XA, one audio sector per 8, complete 2352-byte sectors at their
    LBA slots of the file image. Many independent files per step: their video (FDCT/pack/framing
    kernels) and XA chains (ADPCM/framing kernels) run concurrently on two streams into the same
    images. Also timed: each half alone, and ONE long file (bound by its two serial XA chains)."""
    torch, pb = ctx.torch, ctx.pb
    from psxavenc_b200 import synth
    w, h, fpf, interleave = 320, 240, 8, 8
    files = FRAMES_PER_STEP // fpf
    n = files * fpf
    wl = Workload("strcd", w, h, 0, 18144, n, noise=3)
    params = pb.str_params(pb.FORMAT_STRCD, 150 * (interleave - 1), 15 * interleave, framing=1, interleave=interleave,
                           place_at_lba=1, xa_file=1, xa_channel=0, frames_per_file=fpf)
    samples = 10 * 2016                       # 8 frames at 15 fps = 0.533 s = 20160 sample frames = 10 XA sectors
    image_bytes = int(pb.lib().psxb200_strcd_image_bytes(C.byref(params), fpf, 4, 1, samples))
    params.file_stride = image_bytes
    host = wl.host_frames(ctx.rank * n)
    d_frames = torch.from_numpy(host).to(ctx.dev)
    pcm_one = np.concatenate([synth.gen_pcm(samples, 2, 5 + ctx.rank).ravel(), np.zeros(256, np.int16)])
    pcm_stride = len(pcm_one)
    d_pcm = torch.from_numpy(pcm_one).to(ctx.dev).unsqueeze(0).repeat(files, 1).contiguous()
    gen = torch.Generator(device=ctx.dev)
    gen.manual_seed(77 + ctx.rank)
    d_pcm[1:, :2 * samples] += torch.randint(-32, 32, (files - 1, 2 * samples), dtype=torch.int16, device=ctx.dev, generator=gen)
    d_images = torch.zeros((files, image_bytes), dtype=torch.uint8, device=ctx.dev)
    d_res = torch.zeros((n, 4), dtype=torch.int32, device=ctx.dev)
    d_states = torch.zeros((files, 48), dtype=torch.uint8, device=ctx.dev)
    enc = pb.BsEncoder(0, w, h, ctx.fdct, max_batch=n)
    lib = pb.lib()
    side = torch.cuda.Stream()
    fork, join = torch.cuda.Event(), torch.cuda.Event()

    def video(stream):
        rc = lib.psxb200_str_encode_device_ex(enc.handle, n, d_frames.data_ptr(), C.byref(params), d_images.data_ptr(),
                                              d_res.data_ptr(), stream.cuda_stream)
        assert rc == 0, pb.last_error()

    def audio(stream):
        d_states.zero_() if stream is ctx.stream else None
        rc = lib.psxb200_xa_encode_device_ex(files, 1, 1, 37800, 4, 1, 0, d_pcm.data_ptr(), pcm_stride, samples, 0, interleave,
                                             d_states.data_ptr(), d_images.data_ptr(), image_bytes, interleave * 2352,
                                             stream.cuda_stream)
        assert rc > 0, pb.last_error()

    def both():
        d_images.zero_()
        d_states.zero_()
        fork.record(ctx.stream)
        side.wait_event(fork)
        audio(side)
        video(ctx.stream)
        join.record(side)
        ctx.stream.wait_event(join)

    steps = max(5, min(ctx.args.steps, 20))
    for _ in range(3):
        both()
    ms_both, p_both = ctx.timed(both, steps)
    ms_video, p_video = ctx.timed(lambda: video(ctx.stream), steps)
    ms_audio, p_audio = ctx.timed(lambda: audio(ctx.stream), steps)
    both()
    torch.cuda.synchronize()
    ms_both, ms_video, ms_audio = ctx.max_over_ranks([ms_both / p_both, ms_video / p_video, ms_audio / p_audio])

    # parity on every rank: three of its files against the reference's encode_file_str loop
    backend = cpu_backend()
    ok = True
    checked = 0
    if hasattr(backend, "str_mux"):
        for f in (0, 1, files - 1):
            pcm_f = d_pcm[f].cpu().numpy()
            exp, _ = backend.str_mux(0, w, h, d_frames[f * fpf:(f + 1) * fpf].cpu().numpy(), fmt=pb.FORMAT_STRCD, pcm=pcm_f,
                                     n_samples=samples, fdct=ctx.fdct)
            ok = ok and np.array_equal(d_images[f].cpu().numpy(), exp.ravel())
            checked += 1
    if not ctx.all_ranks_ok(ok):
        raise SystemExit("bench.py: strcd file image differs from the reference's encode_file_str loop")

    out = {"workload": "strcd: %d independent files per GPU per step, each %d frames 320x240 BS v2 (budgets 16128,18144,18144,...) "
                       "+ %d XA sample frames (37800 Hz 4-bit stereo), complete 2352-byte sectors, 1 audio per 8" % (files, fpf, samples),
           "value": ctx.world * n / (ms_both / 1000.0), "unit": UNIT, "ms_per_step": ms_both,
           "video_alone_ms": ms_video, "audio_alone_ms": ms_audio,
           "overlap": "video and XA on two streams: %.3f ms together vs %.3f ms back to back" % (ms_both, ms_video + ms_audio),
           "audio_msamples_per_sec": ctx.world * files * samples * 2 / (ms_audio / 1000.0) / 1e6,
           "files_checked_per_rank": checked,
           "algorithmic_bytes_per_frame": 115200 + image_bytes // fpf}

    # one long file: 1024 frames with their 68 s of audio — two serial XA chains set the pace
    if ctx.rank == 0:
        long_frames = 1024
        p1 = pb.str_params(pb.FORMAT_STRCD, 150 * 7, 15 * 8, framing=1, interleave=8, place_at_lba=1, xa_file=1, xa_channel=0)
        lsamples = long_frames * 2520
        lbytes = int(lib.psxb200_strcd_image_bytes(C.byref(p1), long_frames, 4, 1, lsamples))
        l_image = torch.zeros(lbytes, dtype=torch.uint8, device=ctx.dev)
        l_pcm = d_pcm[:, :2 * samples].reshape(-1)[:2 * lsamples + 512].contiguous()
        l_states = torch.zeros(48, dtype=torch.uint8, device=ctx.dev)
        l_res = torch.zeros((long_frames, 4), dtype=torch.int32, device=ctx.dev)
        torch.cuda.synchronize()
        t = []
        for what in ("both", "audio"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ctx.stream)
            if what == "both":
                fork.record(ctx.stream)
                side.wait_event(fork)
            s_audio = side if what == "both" else ctx.stream
            l_states.zero_()
            rc = lib.psxb200_xa_encode_device_ex(1, 1, 1, 37800, 4, 1, 0, l_pcm.data_ptr(), 0, lsamples, 0, 8, l_states.data_ptr(),
                                                 l_image.data_ptr(), 0, 8 * 2352, s_audio.cuda_stream)
            assert rc > 0, pb.last_error()
            if what == "both":
                rc = lib.psxb200_str_encode_device_ex(enc.handle, long_frames, d_frames.data_ptr(), C.byref(p1), l_image.data_ptr(),
                                                      l_res.data_ptr(), ctx.stream.cuda_stream)
                assert rc == 0, pb.last_error()
                join.record(side)
                ctx.stream.wait_event(join)
            b.record(ctx.stream)
            torch.cuda.synchronize()
            t.append(a.elapsed_time(b))
        out["single_file"] = {"frames": long_frames, "xa_sample_frames": lsamples, "ms": t[0], "audio_chain_ms": t[1],
                              "value": long_frames / (t[0] / 1000.0), "unit": UNIT,
                              "note": "Amdahl: one file is bound by its two serial XA chains (adpcm.c:135-136,186-190); "
                                      "throughput comes from independent files"}

    # e2e: host buffers through psxb200_strcd_encode_host
    h_frames = torch.empty((n, wl.frame_bytes), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(d_frames)
    h_pcm = torch.empty((files, pcm_stride), dtype=torch.int16, pin_memory=True)
    h_pcm.copy_(d_pcm)
    h_images = torch.empty((files, image_bytes), dtype=torch.uint8, pin_memory=True)
    h_res = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
    host_enc = pb.BsEncoder(0, w, h, ctx.fdct, max_batch=256)

    def host_call():
        rc = lib.psxb200_strcd_encode_host(host_enc.handle, files, fpf, h_frames.data_ptr(), C.byref(params), 37800, 4, 1,
                                           h_pcm.data_ptr(), pcm_stride, samples, None, h_images.data_ptr(), image_bytes,
                                           h_res.data_ptr())
        assert rc == 0, pb.last_error()

    host_call()
    ctx.barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        host_call()
    dt = time.perf_counter() - t0
    ok = bool(torch.equal(h_images, d_images.cpu()))
    if not ctx.all_ranks_ok(ok):
        raise SystemExit("bench.py: strcd e2e images differ from the device-resident images")
    (dt,) = ctx.max_over_ranks([dt])
    out["e2e"] = {"value": ctx.world * n * reps / dt, "unit": UNIT, "h2d_bytes_per_step": n * wl.frame_bytes + files * pcm_stride * 2,
                  "d2h_bytes_per_step": files * image_bytes + 16 * n, "api": "psxb200_strcd_encode_host (pinned host buffers)"}
    if ctx.rank == 0 and ctx.world == 1 and not ctx.args.no_cpu and hasattr(backend, "str_mux"):
        # the reference's loop on all cores: one file per thread
        cores = os.cpu_count() or 1
        frames_np, pcm_np = d_frames.cpu().numpy(), d_pcm.cpu().numpy()
        count = [0]

        def work(t):
            for f in range(t, 4 * cores, cores):
                backend.str_mux(0, w, h, frames_np[f * fpf:(f + 1) * fpf], fmt=pb.FORMAT_STRCD, pcm=pcm_np[f], n_samples=samples, fdct=ctx.fdct)
                count[0] += 1

        threads = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        cdt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 4 * cores * fpf / cdt, "unit": UNIT, "cores": cores, "kind": backend.kind,
                               "sample": "%d files (video + XA, the reference's encode_file_str loop), one per thread" % (4 * cores)}
    host_enc.close()
    enc.close()
    return out


def bench_vagi(ctx, files):
    """`vagi` (x files): 8-channel 44.1 kHz SPU-ADPCM, 2048-byte interleave = 3584 samples per channel
    per chunk (filefmt.c:296, 319-341). files = 1: one long stream (8 serial chains, latency-bound);
    files = 1024: independent streams (throughput)."""
    torch, pb = ctx.torch, ctx.pb
    from psxavenc_b200 import synth
    import oracle
    ch = 8
    count = 3584 * 4 if files > 1 else 3584 * 256          # 1 file: 20.8 s of audio per channel
    base = synth.gen_pcm(count, ch, 7 + ctx.rank)
    d_pcm = torch.from_numpy(base).to(ctx.dev).unsqueeze(0).repeat(files, 1, 1).contiguous()
    if files > 1:
        gen = torch.Generator(device=ctx.dev)
        gen.manual_seed(99 + ctx.rank)
        d_pcm[1:] += torch.randint(-64, 64, d_pcm[1:].shape, dtype=torch.int16, device=ctx.dev, generator=gen)
    streams = files * ch
    row = 16 * (count // 28)
    d_states = torch.zeros((streams, 24), dtype=torch.uint8, device=ctx.dev)
    d_out = torch.zeros((streams, row), dtype=torch.uint8, device=ctx.dev)
    lib = pb.lib()

    def step():
        d_states.zero_()
        rc = lib.psxb200_spu_encode_device(streams, d_pcm.data_ptr(), ch, count * ch, count, None, d_states.data_ptr(),
                                           d_out.data_ptr(), row, ctx.stream.cuda_stream)
        assert rc == 0, pb.last_error()

    for _ in range(3):
        step()
    steps = max(3, min(ctx.args.steps, 10)) if files > 1 else 3
    ms, passes = ctx.timed(step, steps)
    (ms,) = ctx.max_over_ranks([ms])
    value = ctx.world * streams * count * passes / (ms / 1000.0) / 1e6
    bytes_per_sample = 2.0 + 16.0 / 28.0
    gbs = value * 1e6 * bytes_per_sample / 1e9
    out = {"metric": "spu_adpcm_msamples_per_sec", "value": value, "unit": "Msamples/s",
           "workload": "vagi x %d: %d independent 8-channel streams per GPU, %d samples per channel" % (files, files, count),
           "ms_per_step": ms / passes,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": ctx.peak, "unit": "GB/s", "frac": gbs / ctx.peak,
                        "algorithmic_bytes_per_sample": bytes_per_sample,
                        "traffic": load_traffic().get("adpcm_spu_kernel") if files == 1024 else None,
                        "note": "ALU/issue bound by design (SURVEY.md 8d: ~375 integer ops per sample, ceiling ~1e5 Msamples/s "
                                "per GPU with >= 1e4 chains); one chain runs 28 samples per ~1.3 us",
                        "frac_of_alu_ceiling": value / ctx.world / 1.0e5}}
    # parity on EVERY rank: all chains of file 0 and of the last file against the CPU reference
    backend = cpu_backend()
    got_all = d_out.cpu().numpy()
    ok = True
    cpu_dt = 0.0
    for f in sorted({0, files - 1}):
        pcm_f = d_pcm[f].cpu().numpy()
        outs = [None] * ch

        def work(c):
            st = oracle.ChannelState()
            outs[c] = backend.spu_encode(st, pcm_f, count, ch, offset=c)

        threads = [threading.Thread(target=work, args=(c,)) for c in range(ch)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        cpu_dt += time.perf_counter() - t0
        ok = ok and all(np.array_equal(got_all[f * ch + c], outs[c]) for c in range(ch))
    if not ctx.all_ranks_ok(ok):
        raise SystemExit("bench.py: SPU-ADPCM GPU output differs from the CPU reference on some rank")
    out["parity"] = "all 8 chains of file 0 and file %d on every rank" % (files - 1)
    checked_files = len({0, files - 1})
    if ctx.rank == 0 and ctx.world == 1 and not ctx.args.no_cpu:
        out["cpu_baseline"] = {"value": checked_files * ch * count / cpu_dt / 1e6, "unit": "Msamples/s",
                               "cores": min(os.cpu_count() or 1, ch), "kind": backend.kind,
                               "sample": "%d file(s) x 8 channels x %d samples, one thread per channel" % (checked_files, count)}

    # e2e: host buffers through psxb200_spu_encode_host
    h_pcm = torch.empty(d_pcm.shape, dtype=torch.int16, pin_memory=True)
    h_pcm.copy_(d_pcm)
    h_out = torch.empty((streams, row), dtype=torch.uint8, pin_memory=True)
    h_states = torch.zeros((streams, 24), dtype=torch.uint8).pin_memory()

    def host_call():
        h_states.zero_()
        rc = lib.psxb200_spu_encode_host(streams, h_pcm.data_ptr(), ch, count * ch, count, h_states.data_ptr(), h_out.data_ptr(), row)
        assert rc == 0, pb.last_error()

    host_call()
    ctx.barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        host_call()
    dt = time.perf_counter() - t0
    if not ctx.all_ranks_ok(bool(np.array_equal(h_out.numpy(), got_all))):
        raise SystemExit("bench.py: SPU-ADPCM e2e output differs from the device-resident output")
    (dt,) = ctx.max_over_ranks([dt])
    out["e2e"] = {"value": ctx.world * streams * count * reps / dt / 1e6, "unit": "Msamples/s",
                  "h2d_bytes_per_step": streams * count * 2 + streams * 24, "d2h_bytes_per_step": streams * row + streams * 24,
                  "api": "psxb200_spu_encode_host (pinned host buffers)"}
    # the same file(s) of rank 0 dealt over ALL GPUs by one process: whole files per device when
    # there are enough of them, else channel c on device c mod G — a single stream
    # gains nothing from more devices, its channels are serial chains that already run side by side
    if ctx.world > 1 and pb.device_count() >= ctx.world:
        ctx.host_barrier()
        if ctx.rank == 0:
            m_out = torch.empty((streams, row), dtype=torch.uint8, pin_memory=True)

            def multi_call():
                h_states.zero_()
                rc = lib.psxb200_spu_encode_host_multi(ctx.world, None, streams, h_pcm.data_ptr(), ch, count * ch, count,
                                                       h_states.data_ptr(), m_out.data_ptr(), row)
                assert rc == 0, pb.last_error()

            multi_call()
            t0 = time.perf_counter()
            for _ in range(reps):
                multi_call()
            mdt = time.perf_counter() - t0
            if not np.array_equal(m_out.numpy(), got_all):
                raise SystemExit("bench.py: SPU-ADPCM output of the multi-device entry differs")
            out["over_devices"] = {"value": streams * count * reps / mdt / 1e6, "unit": "Msamples/s", "devices": ctx.world,
                                   "split": "whole files per device" if files >= ctx.world else "channel c on device c mod G",
                                   "api": "psxb200_spu_encode_host_multi (one process, rank 0's streams only)"}
        ctx.host_barrier()
    return out


def bench_spu_sine(ctx):
    """BASELINE configs[0] `spu`: mono 22050 Hz 440 Hz sine, 60 s, raw SPU-ADPCM, driven the way
    encode_file_spu does (filefmt.c:212-292): ONE <= 28-sample block per psx_audio_spu_encode call.
    The reference's CPU code is timed beside the drop-in symbol on the same calls; the config is
    "reference plumbing, no GPU" — a serial chain fed one block at a time is latency-bound on a GPU."""
    pb = ctx.pb
    from psxavenc_b200 import synth
    backend = cpu_backend()
    seconds = 60
    pcm = synth.gen_sine(22050 * seconds)
    blocks = (len(pcm) + 27) // 28

    def drive(lib, limit_blocks):
        fn = lib.psx_audio_spu_encode
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        state = pb.ChannelState()
        out = np.zeros((limit_blocks, 16), np.uint8)
        sp, base, op = C.addressof(state), pcm.ctypes.data, out.ctypes.data
        t0 = time.perf_counter()
        for b in range(limit_blocks):
            fn(sp, base + 56 * b, min(28, len(pcm) - 28 * b), 1, op + 16 * b)
        return time.perf_counter() - t0, out

    gpu_blocks = min(blocks, 20000)                      # bounded: ~12 s of the 60 s through the GPU drop-in
    drive(pb.lib(), 64)
    gpu_dt, gpu_out = drive(pb.lib(), gpu_blocks)
    out = {"workload": "spu: mono 22050 Hz 440 Hz sine, %d s, one <=28-sample block per psx_audio_spu_encode call" % seconds,
           "dropin_gpu": {"value": gpu_blocks * 28 / gpu_dt / 1e6, "unit": "Msamples/s", "us_per_call": gpu_dt / gpu_blocks * 1e6,
                          "blocks": gpu_blocks, "api": "psx_audio_spu_encode (drop-in symbol; samples in the kernel parameters, results and completion flag in mapped host memory)"}}
    if hasattr(backend, "lib") and backend.kind == "reference":
        cpu_dt, cpu_out = drive(backend.lib, blocks)
        if not np.array_equal(cpu_out[:gpu_blocks], gpu_out):
            raise SystemExit("bench.py: spu sine output of the drop-in differs from the reference")
        out["cpu_reference"] = {"value": blocks * 28 / cpu_dt / 1e6, "unit": "Msamples/s", "us_per_call": cpu_dt / blocks * 1e6,
                                "blocks": blocks, "cores": 1, "kind": "reference"}
        out["parity_blocks_checked"] = gpu_blocks
    # the whole sine as ONE call (what a batching caller would do): still one serial chain
    st = pb.ChannelState()
    whole = np.zeros(16 * blocks, np.uint8)
    t0 = time.perf_counter()
    pb.lib().psx_audio_spu_encode(C.addressof(st), pcm.ctypes.data, len(pcm), 1, whole.ctypes.data)
    one_dt = time.perf_counter() - t0
    if not np.array_equal(whole[:16 * gpu_blocks].reshape(-1, 16), gpu_out):
        raise SystemExit("bench.py: spu sine, single call differs from per-block calls")
    out["single_call_gpu"] = {"value": len(pcm) / one_dt / 1e6, "unit": "Msamples/s", "ms": one_dt * 1e3}
    return out


def run_ours(args):
    ctx = Ctx(args)
    line = {}
    bench_headline(ctx, line)

    configs = {}
    legs = [("strv_easy", lambda: bench_content(ctx, 0, "easy")), ("strv_hard", lambda: bench_content(ctx, 6, "hard")),
            ("sbs", lambda: bench_sbs(ctx)), ("strcd", lambda: bench_strcd(ctx)),
            ("vagi_x1024", lambda: bench_vagi(ctx, 1024)), ("vagi_x1", lambda: bench_vagi(ctx, 1)),
            # the rest of SURVEY.md 8d's B set: x64 leaves most SMs idle, x4096 is the kernel with the machine full
            ("vagi_x64", lambda: bench_vagi(ctx, 64)), ("vagi_x4096", lambda: bench_vagi(ctx, 4096))]
    if args.only:
        legs = [l for l in legs if l[0] in args.only.split(",")]
    for name, fn in legs:
        if args.headline_only:
            break
        try:
            configs[name] = fn()
        except SystemExit:
            raise
        except Exception as e:      # an auxiliary leg must not take the headline down with it
            import traceback
            traceback.print_exc()
            configs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        ctx.torch.cuda.empty_cache()
    if ctx.rank == 0 and not args.headline_only and (not args.only or "spu" in args.only.split(",")):
        try:
            configs["spu"] = bench_spu_sine(ctx)
        except SystemExit:
            raise
        except Exception as e:
            configs["spu"] = {"error": "%s: %s" % (type(e).__name__, e)}
    ctx.barrier()
    if ctx.rank == 0:
        line["configs"] = configs
        if "vagi_x1024" in configs:
            line["adpcm"] = configs["vagi_x1024"]      # second half of BASELINE.json's metric (SPU-ADPCM Msamples/s)
        line["gpu_launches_total"] = int(ctx.pb.launch_count())
        emit(line)
    if ctx.dist:
        ctx.dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout. Libraries underneath (NCCL's version banner,
    anything printf'ing from C) write to file descriptor 1 as well, so fd 1 is pointed at stderr
    for the whole run and the JSON line alone goes to the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fdct", default="sse2", choices=["sse2", "islow"],
                    help="which FFmpeg AVDCT.fdct both arms reproduce bit-exactly (sse2 = this box's libavcodec)")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("PSXB200_CHUNK", str(FRAMES_PER_STEP))),
                    help="frames per internal kernel launch (device-resident path)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--headline-only", action="store_true", help="strv only, skip the other configs (profiling)")
    ap.add_argument("--only", default="", help="comma-separated config legs to run beside the headline")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
