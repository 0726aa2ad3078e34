#!/usr/bin/env python3
"""Benchmark of the B200 MDEC/BS encode core (BASELINE.json config `strv`).

A "step" encodes one batch of 4096 synthetic 320x240 NV21 frames per GPU to BS v2 bitstreams
with a 20160-byte budget per frame (psxavenc `-t strv` defaults: 15 fps at 2x CD speed = 10
sectors of 2016 bytes per frame, filefmt.c:540-552; mdec.c:772-774) through the C ABI of
libpsxav_b200.so. One JSON line is printed by rank 0:

  value      frames/s, whole job, inputs resident in HBM, timed with CUDA events on the
             launching stream, max over ranks
  e2e        the same metric through psxb200_bs_encode_host with pinned HOST buffers
             (host->device and device->host copies inside the timed region)
  roofline   dominant kernel: algorithmic bytes per launch / CUDA-event launch duration
             against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
             the UNMODIFIED reference C (oracle/_ref/libpsxav_ref.so; falls back to the
             oracle port when that cannot be loaded) on all host cores, bounded sample

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 320, 240
CODEC_V2 = 0
FRAME_MAX_SIZE = 20160
FRAME_BYTES = WIDTH * HEIGHT * 3 // 2
FRAMES_PER_STEP = 4096            # per GPU (weak scaling)
NOISE_BITS = 3                    # "typical" content: the reference settles on quant scale 2
ALGO_BYTES_PER_FRAME = FRAME_BYTES + FRAME_MAX_SIZE      # 135 360 B
METRIC = "bs_v2_320x240_frames_per_sec"
UNIT = "frames/s"
WORKLOAD = ("strv: 320x240 BS v2, frame_max_size 20160 B, %d synthetic NV21 frames per step per GPU, "
            "noise_bits=%d (quant scale 2)" % (FRAMES_PER_STEP, NOISE_BITS))
SMOOTH = False


def select_workload(name, noise):
    """The headline is `strv` (BASELINE.json configs[1]); the other shapes exist for profiling
    and DESIGN.md tables only and are never what the driver's default invocation measures."""
    global WIDTH, HEIGHT, CODEC_V2, FRAME_MAX_SIZE, FRAME_BYTES, FRAMES_PER_STEP, NOISE_BITS
    global ALGO_BYTES_PER_FRAME, METRIC, WORKLOAD, SMOOTH
    if name == "sbs":          # BASELINE.json configs[4]: 640x480 BS v3, 8192-byte frames
        WIDTH, HEIGHT, CODEC_V2, FRAME_MAX_SIZE, FRAMES_PER_STEP, SMOOTH = 640, 480, 1, 8192, 1024, True
        METRIC = "bs_v3_640x480_frames_per_sec"
        WORKLOAD = "sbs: 640x480 BS v3, frame_max_size 8192 B, %d smooth synthetic frames per step per GPU" % FRAMES_PER_STEP
    elif noise != NOISE_BITS:
        NOISE_BITS = noise
        WORKLOAD = ("strv: 320x240 BS v2, frame_max_size 20160 B, %d synthetic NV21 frames per step per GPU, "
                    "noise_bits=%d" % (FRAMES_PER_STEP, NOISE_BITS))
    FRAME_BYTES = WIDTH * HEIGHT * 3 // 2
    ALGO_BYTES_PER_FRAME = FRAME_BYTES + FRAME_MAX_SIZE


def fdct_from_name(name):
    return 1 if name == "sse2" else 0


def make_frames(count, first, distinct=256):
    """`distinct` integer-generator frames tiled to `count`; the copies
    are made unique on the device by the caller."""
    from psxavenc_b200 import synth
    if SMOOTH:
        base = np.stack([synth.gen_smooth_frame(first + i, WIDTH, HEIGHT) for i in range(min(distinct // 4, count))])
    else:
        base = synth.gen_frames(first, min(distinct, count), WIDTH, HEIGHT, NOISE_BITS)
    reps = (count + len(base) - 1) // len(base)
    return np.tile(base, (reps, 1))[:count]


# ---------------------------------------------------------------------------------------
# CPU side: the reference (or the oracle port) on all host cores
# ---------------------------------------------------------------------------------------

class CpuEncoder:
    """Runs the CPU implementation of the path over slices of a frame batch on `cores` threads
    (ctypes releases the GIL; every thread owns its encoder handle)."""

    def __init__(self, fdct):
        import oracle
        self.fdct = fdct
        self.cores = os.cpu_count() or 1
        try:
            self.backend = oracle.Reference()
        except Exception:
            self.backend = oracle.Restated()
        self.kind = self.backend.kind

    def encode(self, frames):
        n = len(frames)
        chunks = [c for c in np.array_split(np.arange(n), min(self.cores, n)) if len(c)]
        results = [None] * len(chunks)

        def work(i):
            sl = frames[chunks[i][0]:chunks[i][-1] + 1]
            results[i] = self.backend.bs_encode_batch(CODEC_V2, WIDTH, HEIGHT, sl, FRAME_MAX_SIZE, self.fdct)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(chunks))]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        out = np.concatenate([r[0] for r in results])
        res = np.concatenate([r[1] for r in results])
        return dt, out, res

    def calibrate(self, frames, target_seconds):
        """Frames per bounded sample so that one sample costs about target_seconds."""
        probe = frames[:max(self.cores, 8)]
        dt, _, _ = self.encode(probe)
        rate = len(probe) / dt
        n = int(rate * target_seconds)
        return max(self.cores, min(len(frames), n // self.cores * self.cores))


def cpu_baseline(frames, fdct, target_seconds=1.5):
    """All host cores on the step batch, repeated until about target_seconds of wall time
    (= target_seconds x cores of CPU work, i.e. 10-30 core-seconds on the usual boxes)."""
    cpu = CpuEncoder(fdct)
    total, reps, out, res = 0.0, 0, None, None
    while total < target_seconds and reps < 64:
        dt, out, res = cpu.encode(frames)
        total += dt
        reps += 1
    return {
        "value": reps * len(frames) / total, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
        "sample": "the %d frames of the step batch x %d passes, %d threads, %.1f s wall" % (len(frames), reps, cpu.cores, total),
    }, out, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fdct = fdct_from_name(args.fdct)
    cpu = CpuEncoder(fdct)
    frames = make_frames(FRAMES_PER_STEP, 0)
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
        "config": {"workload": WORKLOAD, "fdct": args.fdct, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML (in-process, every few
    milliseconds) while the timed region runs."""
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
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def run_ours(args):
    import torch
    import psxavenc_b200 as pb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or pb.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    fdct = fdct_from_name(args.fdct)
    n = FRAMES_PER_STEP
    dev = torch.device("cuda", local)

    # ---- inputs: resident in HBM, every frame distinct, 472 MB per GPU (> 126 MB L2) ------
    host_frames = make_frames(n, rank * n)
    d_frames = torch.from_numpy(host_frames).to(dev)
    if not SMOOTH:
        # flip the lowest luma bit of the tiled copies so that no two frames are identical
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        salt = torch.randint(0, 2, (n, WIDTH * HEIGHT), dtype=torch.uint8, device=dev, generator=gen)
        salt[:256] = 0
        d_frames[:, :WIDTH * HEIGHT] ^= salt
        del salt
    d_sizes = torch.full((n,), FRAME_MAX_SIZE, dtype=torch.int32, device=dev)
    d_out = torch.zeros((n, FRAME_MAX_SIZE), dtype=torch.uint8, device=dev)
    d_res = torch.zeros((n, 4), dtype=torch.int32, device=dev)
    gathered = torch.zeros((world * n, 4), dtype=torch.int32, device=dev) if world > 1 else None

    enc = pb.BsEncoder(CODEC_V2, WIDTH, HEIGHT, fdct, max_batch=args.chunk)
    enc.timing(True)
    stream = torch.cuda.current_stream()

    def step():
        enc.encode_device(n, d_frames, d_sizes, FRAME_MAX_SIZE, d_out, FRAME_MAX_SIZE, d_res, stream.cuda_stream)
        if world > 1:
            # the path's only exchange: per-frame {bytes_used, blocks_used, q, hwords} to the muxing rank
            dist.all_gather_into_tensor(gathered, d_res)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    enc.read_timing()

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = pb.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for _ in range(args.steps):
        step()
    t_end.record(stream)
    barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    launches = pb.launch_count() - launches0
    dct_ms, pack_ms, pairs = enc.read_timing()
    clocks = sampler.stop() if sampler else None
    enc.timing(False)

    # ---- parity spot check against the CPU oracle on this run's own bytes -------------------
    res = d_res.cpu().numpy()
    parity = None
    if rank == 0:
        import oracle
        sel = np.array([0, 1, 255, 256, 257, min(1000, n - 2), n - 1])
        frames_sel = d_frames[torch.from_numpy(sel).to(dev)].cpu().numpy()
        exp_out, exp_res = oracle.Restated().bs_encode_batch(CODEC_V2, WIDTH, HEIGHT, frames_sel, FRAME_MAX_SIZE, fdct)
        got_out = d_out[torch.from_numpy(sel).to(dev)].cpu().numpy()
        parity = bool(np.array_equal(got_out, exp_out) and np.array_equal(res[sel], exp_res))
        if not parity:
            raise SystemExit("bench.py: GPU output differs from the CPU oracle — number withheld")

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ------------------
    h_frames = torch.empty((n, FRAME_BYTES), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(d_frames)
    h_sizes = torch.full((n,), FRAME_MAX_SIZE, dtype=torch.int32).pin_memory()
    h_out = torch.empty((n, FRAME_MAX_SIZE), dtype=torch.uint8, pin_memory=True)
    h_res = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        enc.encode_host_into(n, h_frames, h_sizes, h_out, FRAME_MAX_SIZE, h_res)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        failed = enc.encode_host_into(n, h_frames, h_sizes, h_out, FRAME_MAX_SIZE, h_res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert failed == 0
    if rank == 0:
        assert np.array_equal(h_res.numpy(), res) and torch.equal(h_out, d_out.cpu()), "e2e output differs"

    # the same host->device link measured bare (one pinned cudaMemcpy of the step's frames): the
    # e2e path is bound by it, so its share of this figure is what the pipeline can be judged by
    d_probe = torch.empty_like(d_frames)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_probe.copy_(h_frames, non_blocking=True)
    a.record(stream)
    for _ in range(3):
        d_probe.copy_(h_frames, non_blocking=True)
    b.record(stream)
    torch.cuda.synchronize()
    link_gbs = 3 * n * FRAME_BYTES / (a.elapsed_time(b) / 1000.0) / 1e9
    del d_probe

    times = torch.tensor([elapsed_ms, e2e_s * 1000.0, dct_ms, pack_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, dct_ms, pack_ms = times.tolist()

    # ---- SPU-ADPCM (second half of BASELINE.json's metric), reported beside the headline ---------
    adpcm = bench_adpcm(pb, torch, dev, stream, rank, world, dist, args)

    if rank == 0:
        value = world * n * args.steps / (elapsed_ms / 1000.0)
        peak, peak_src = load_peak()
        frames_per_launch = min(args.chunk, n)
        dominant = "bs_pack_kernel" if pack_ms >= dct_ms else "bs_dct_kernel"
        dom_ms = max(pack_ms, dct_ms) / max(pairs, 1)
        achieved = ALGO_BYTES_PER_FRAME * frames_per_launch / (dom_ms / 1000.0) / 1e9
        traffic = load_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "fdct": args.fdct, "frames_per_launch": frames_per_launch,
                       "l2": "inputs larger than L2 (%.0f MB of frames per step per GPU)" % (n * FRAME_BYTES / 1e6),
                       "quant_scale_mean": float(res[:, 2].mean()), "parity_spot_check": parity,
                       "collective": "all_gather of per-frame results" if world > 1 else "none"},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic.get(dominant),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME * frames_per_launch,
                         "launch_ms": dom_ms, "kernel_share": {"bs_dct_kernel": dct_ms / (dct_ms + pack_ms),
                                                               "bs_pack_kernel": pack_ms / (dct_ms + pack_ms)},
                         "kernel_ms_total": dct_ms + pack_ms, "step_ms_total": elapsed_ms},
            "e2e": {"value": world * n * e2e_steps / (e2e_ms / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": n * (FRAME_BYTES + 4), "d2h_bytes_per_step": n * (FRAME_MAX_SIZE + 16),
                    "steps": e2e_steps, "timer": "host wall clock around the synchronous C-ABI call, max over ranks",
                    "h2d_gbs_per_gpu": n * (FRAME_BYTES + 4) * e2e_steps / (e2e_ms / 1000.0) / 1e9,
                    "h2d_link_gbs_rank0": link_gbs,
                    "api": "psxb200_bs_encode_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "adpcm": adpcm,
        }
        if world == 1 and not args.no_cpu:
            base, cpu_out, cpu_res = cpu_baseline(d_frames.cpu().numpy(), fdct)
            k = len(cpu_res)
            if not (np.array_equal(cpu_res, res[:k]) and np.array_equal(cpu_out, d_out[:k].cpu().numpy())):
                raise SystemExit("bench.py: GPU output differs from the CPU baseline's output")
            base["parity_frames_checked"] = k
            line["cpu_baseline"] = base
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_adpcm(pb, torch, dev, stream, rank, world, dist, args):
    """`vagi x B`: B independent 8-channel 44.1 kHz streams, 4 interleave chunks of 3584 samples
    per channel each (filefmt.c:296, 319-341), SPU-ADPCM through psxb200_spu_encode_device."""
    from psxavenc_b200 import synth
    files, ch, count = 1024, 8, 3584 * 4
    base = synth.gen_pcm(count, ch, 7 + rank)
    d_pcm = torch.from_numpy(base).to(dev).unsqueeze(0).repeat(files, 1, 1).contiguous()
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    d_pcm += torch.randint(-64, 64, d_pcm.shape, dtype=torch.int16, device=dev, generator=gen)
    streams = files * ch
    row = 16 * (count // 28)
    d_states = torch.zeros((streams, 24), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((streams, row), dtype=torch.uint8, device=dev)

    def step():
        d_states.zero_()
        rc = pb.lib().psxb200_spu_encode_device(streams, d_pcm.data_ptr(), ch, count * ch, count, None,
                                                d_states.data_ptr(), d_out.data_ptr(), row, stream.cuda_stream)
        assert rc == 0, pb.last_error()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    steps = max(3, min(args.steps, 10))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        step()
    b.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    samples = world * streams * count * steps
    result = {"metric": "spu_adpcm_msamples_per_sec", "value": samples / (ms.item() / 1000.0) / 1e6, "unit": "Msamples/s",
              "workload": "vagi x %d: %d independent 8-channel streams per GPU, %d samples per channel" % (files, files, count),
              "algorithmic_bytes_per_sample": 2.0 + 16.0 / 28.0}
    result["hbm_gbs"] = result["value"] * 1e6 * result["algorithmic_bytes_per_sample"] / 1e9
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        try:
            cpu = oracle.Reference()
        except Exception:
            cpu = oracle.Restated()
        pcm0 = d_pcm[0].cpu().numpy()
        got = d_out[:ch].cpu().numpy()
        cores = os.cpu_count() or 1
        outs = [None] * ch

        def work(c):
            st = oracle.ChannelState()
            outs[c] = cpu.spu_encode(st, pcm0, count, ch, offset=c)

        t0 = time.perf_counter()
        reps = 4
        for _ in range(reps):
            threads = [threading.Thread(target=work, args=(c,)) for c in range(ch)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        dt = time.perf_counter() - t0
        if not all(np.array_equal(got[c], outs[c]) for c in range(ch)):
            raise SystemExit("bench.py: SPU-ADPCM GPU output differs from the CPU baseline")
        result["cpu_baseline"] = {"value": reps * ch * count / dt / 1e6, "unit": "Msamples/s", "cores": min(cores, ch),
                                  "kind": cpu.kind, "sample": "file 0 (8 channels x %d samples) x %d" % (count, reps)}
    return result


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="strv", choices=["strv", "sbs"], help="strv is the headline; sbs for profiling")
    ap.add_argument("--noise", type=int, default=NOISE_BITS, help="noise_bits of the synthetic strv frames (0 easy, 3 typical, 6 hard)")
    args = ap.parse_args()
    select_workload(args.workload, args.noise)
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
