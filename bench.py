#!/usr/bin/env python
"""bench.py — whole-encode throughput of the B200 JPEG encode path (BASELINE.json: "encode Mpx/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload image16k|frame4k]

A "step" is one complete encode (K1 colour/DCT/quant -> K2 symbol statistics -> host Huffman table build ->
K3 Huffman pack -> K4 byte stuffing) of one synthetic image per GPU.  Default workload: the 16384x16384 (268 Mpx)
image of BASELINE.json configs[3] -- the configuration the metric's target is quoted on; it fits one GPU.
With N > 1 (torchrun, one rank per GPU) every rank encodes its own image (weak scaling, no data-path collective:
a single image is never split).

value   = Mpx/s with the RGB already resident in HBM (scan left in HBM)
e2e     = Mpx/s through jpgenc_encode_rgb with pinned HOST buffers: H2D of the pixels and D2H of the JPEG inside the
          timed region
roofline= the K1 kernel (fused colour+subsample+DCT+quant+zigzag): 6 algorithmic bytes per padded pixel / its
          CUDA-event duration, against the measured copy bandwidth in MEASURED_PEAKS.json
cpu_baseline / --impl reference = the reference encoder itself (oracle/_ref, built from /root/reference by
          oracle/build_ref.sh) on the box's host cores, on a bounded sample of the same synthetic.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "image16k": (16384, 16384, "single 16384x16384 synthetic image (268 Mpx), whole encode"),
    "frame4k": (3840, 2160, "single 3840x2160 synthetic frame, whole encode"),
    "batch1080p": (1920, 1080, "batch of 1024 synthetic 1920x1080 frames sharded by image across the ranks"),
}
BATCH_FRAMES = 1024
CPU_SAMPLE = (2048, 2048)          # bounded sample of the same generator for the CPU arms (~0.7 s per encode)
if os.environ.get("JPGENC_BENCH_CPU_SAMPLE"):                      # tests use a smaller one
    CPU_SAMPLE = tuple(int(x) for x in os.environ["JPGENC_BENCH_CPU_SAMPLE"].split("x"))
METRIC, UNIT = "encode_throughput", "Mpx/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU while the timed region runs (B200_PROFILING.md recipe).  Sampled through NVML
    inside this process: spawning nvidia-smi every few milliseconds initialises NVML over and over and takes driver locks
    that delay the CUDA calls being timed, on every rank at once.  nvidia-smi is the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(t.strip().isdigit() for t in vis.split(",")) and index < len(vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        return [str(sm), str(self.max_sm)] + ["Active" if r & b else "Not Active" for b in bits]

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [c.strip() for c in out.split(",")] if out else None

    def run(self):
        while not self.stop_flag.is_set():
            try:
                row = self._sample_nvml() if self.nvml else self._sample_smi()
                if row:
                    self.rows.append(row)
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arms: the reference encoder on host cores
# ---------------------------------------------------------------------------------------------------------------
def _ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "jpgEnc_ref")
    return p if os.path.exists(p) else None


def _cpu_sample_file():
    from jpgenc_b200.synth import synth_rgb, write_ppm
    w, h = CPU_SAMPLE
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, f"jpgenc_bench_{w}x{h}_{os.getpid()}.ppm")
    write_ppm(path, synth_rgb(w, h, 0))
    return path


def _run_reference_once(path: str, omp_threads: int) -> float:
    """seconds of Image::writeJPEG ("Encoding duration", src/Image.cpp:975) for one encode by the reference CLI"""
    env = dict(os.environ, OMP_NUM_THREADS=str(omp_threads))
    out = subprocess.run([_ref_binary(), path, path + ".jpg"], capture_output=True, text=True, env=env, check=True).stdout
    m = re.search(r"Encoding duration: (\d+) ms", out)
    return int(m.group(1)) / 1e3


def _run_port_once(rgb) -> float:
    from oracle.pyoracle import Oracle
    o = Oracle()
    t = time.perf_counter()
    o.encode_rgb(rgb)
    return time.perf_counter() - t


def cpu_arm(steps: int, warmup: int):
    """-> (Mpx/s, ms_per_step, dict cpu_baseline)"""
    w, h = CPU_SAMPLE
    mpx = w * h / 1e6
    nproc = os.cpu_count() or 1
    if _ref_binary():
        path = _cpu_sample_file()
        try:
            # the reference parallelises with OpenMP over block rows + 3 std::async channel tasks; on small hosts the
            # OpenMP regions cost more than they save, so take whichever thread setting is faster
            cand = {1: _run_reference_once(path, 1)}
            if nproc > 1:
                cand[nproc] = _run_reference_once(path, nproc)
            omp = min(cand, key=cand.get)
            for _ in range(max(0, warmup - 1)):
                _run_reference_once(path, omp)
            times = [_run_reference_once(path, omp) for _ in range(steps)]
        finally:
            for p in (path, path + ".jpg"):
                if os.path.exists(p):
                    os.remove(p)
        kind, cores = "reference", (3 if omp == 1 else omp)
        sample = (f"reference CLI (oracle/_ref/jpgEnc_ref) writeJPEG time on a {w}x{h} crop-size synthetic (same generator, seed 0), "
                  f"OMP_NUM_THREADS={omp} + its 3 std::async channel tasks, {steps} encodes")
    else:
        from jpgenc_b200.synth import synth_rgb
        rgb = synth_rgb(w, h, 0)
        for _ in range(max(1, warmup)):
            _run_port_once(rgb)
        times = [_run_port_once(rgb) for _ in range(steps)]
        kind, cores = "port", 1
        sample = f"oracle C port (oracle/liboracle.so, oracle/_ref absent) on a {w}x{h} synthetic, single thread, {steps} encodes"
    sec = sum(times) / len(times)
    return mpx / sec, sec * 1e3, {"value": round(mpx / sec, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                 "host_cores_available": nproc}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h, desc = WORKLOADS[args.workload]
    v, ms, cb = cpu_arm(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "cpu_sample": f"{CPU_SAMPLE[0]}x{CPU_SAMPLE[1]} synthetic per step (bounded sample of the workload)"},
            "cpu_baseline": cb,
            "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    from jpgenc_b200.capi import Encoder, pinned_empty, pinned_free

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local)

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w, h, desc = WORKLOADS[args.workload]
    enc = Encoder(local)
    npx = w * h
    padded_px = ((w + 15) // 16 * 16) * ((h + 15) // 16 * 16)
    d_rgb = enc.dev_alloc(npx * 3)
    enc.synth_rgb(d_rgb, w, h, rank)                 # a different image per rank
    enc.bind_device_rgb(d_rgb, w, h)
    enc.synchronize()

    # ---- device-resident: K steps of the whole encode, pixels in HBM, scan left in HBM -------------------
    sampler = ClockSampler(local)
    sampler.start()                                   # samples from the first warm-up step to the end of the timed steps
    for _ in range(args.warmup):
        jpeg_bytes = enc.encode_bound(None)
    barrier()
    launches0 = enc.launch_count()
    s0 = enc.stats()                                  # the library keeps running sums of its per-stage CUDA-event times:
    enc.timer_begin()                                 # read once before and once after, no per-step polling in the loop
    for _ in range(args.steps):
        jpeg_bytes = enc.encode_bound(None)
    ms_total = enc.timer_end()
    s1 = enc.stats()
    n_timed = max(1, s1.timed_encodes - s0.timed_encodes)
    k1_ms = [(s1.sum_ms_k1 - s0.sum_ms_k1) / n_timed]; fwd_ms = [(s1.sum_ms_forward - s0.sum_ms_forward) / n_timed]
    st_ms = [(s1.sum_ms_stats - s0.sum_ms_stats) / n_timed]; en_ms = [(s1.sum_ms_entropy - s0.sum_ms_entropy) / n_timed]
    launches = enc.launch_count() - launches0
    barrier()
    clocks = sampler.summary()
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * npx / 1e6 / (ms_step / 1e3)
    stats = enc.stats()

    # ---- end to end: pinned host pixels in, JPEG bytes out in pinned host memory ---------------------------
    host_rgb, host_ptr = pinned_empty(npx * 3)
    enc.d2h(host_rgb, d_rgb)
    out_cap = int(jpeg_bytes) + 4096
    host_out, out_ptr = pinned_empty(out_cap)
    e2e_warm = min(args.warmup, 2) if npx > 50e6 else args.warmup
    for _ in range(max(1, e2e_warm)):
        n = enc.encode_rgb_into(host_ptr, w, h, out_ptr, out_cap)
    e2e_steps = min(args.steps, 20) if npx > 50e6 else args.steps      # 16 ms per step at 268 Mpx: bounded, still >= 0.3 s
    barrier()
    enc.timer_begin()
    for _ in range(e2e_steps):
        n = enc.encode_rgb_into(host_ptr, w, h, out_ptr, out_cap)
    ms_e2e = max_over_ranks(enc.timer_end()) / e2e_steps
    barrier()
    assert n == jpeg_bytes and host_out[0] == 0xFF and host_out[1] == 0xD8 and host_out[n - 1] == 0xD9
    e2e_value = world * npx / 1e6 / (ms_e2e / 1e3)
    st_e2e = enc.stats()

    # ---- roofline of the dominant kernel (K1) ----------------------------------------------------------------
    peak, peak_src = measured_peak()
    k1 = sum(k1_ms) / len(k1_ms)
    alg_bytes = 6 * padded_px                          # 3 B RGB in + 1.5 samples x int16 out per padded pixel
    achieved = alg_bytes / (k1 / 1e3) / 1e9
    roofline = {"kernel": "forward_kernel (K1: colour+subsample+DCT+quant+zigzag)", "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": round(k1, 4)}
    tr = os.path.join(ROOT, "profiles", "k1_traffic.json")     # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            if t.get("workload") == args.workload:
                roofline["traffic"] = t.get("dram_bytes_per_launch")
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": desc, "width": w, "height": h, "images_per_gpu": 1, "l2": "inputs larger than L2 (805 MB RGB + 805 MB coefficients per step)"
                   if npx > 50e6 else "inputs smaller than L2", "parallelism": f"independent images x{world}",
                   "subsampling": "4:2:0 mean", "tables": "image-optimal length-limited Huffman (host build per image)"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": npx * 3, "d2h_bytes_per_step": int(n),
                "ms_per_step": round(ms_e2e, 3), "steps": e2e_steps, "ms_h2d": round(st_e2e.ms_h2d, 3), "ms_d2h": round(st_e2e.ms_d2h, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "stage_ms": {"k1_forward": round(k1, 4), "k1_plus_refine": round(sum(fwd_ms) / len(fwd_ms), 4),
                     "k2_stats": round(sum(st_ms) / len(st_ms), 4), "k3_k4_entropy": round(sum(en_ms) / len(en_ms), 4)},
        "jpeg_bytes": int(jpeg_bytes), "refined_blocks": int(stats.refined_blocks), "n_blocks": int(stats.n_blocks),
        "whole_encode_gbps_irreducible": round((npx * 3 + jpeg_bytes) / (ms_step / 1e3) / 1e9, 1),
    }
    pinned_free(host_ptr)
    pinned_free(out_ptr)
    enc.dev_free(d_rgb)

    if rank == 0 and world == 1:
        try:
            line["extra"] = extra_workloads(enc, args, peak)
        except Exception as ex:          # extras never invalidate the headline line
            line["extra"] = {"error": repr(ex)}
        _, _, line["cpu_baseline"] = cpu_arm(steps=max(3, min(args.steps, 10)), warmup=1)
    enc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def batch_frames_per_s(local, frames_ptrs, w, h, workers, device_frames, out_ptrs=None, caps=None, reps=1, enc=None):
    """frames/s of one synchronous batch call on this rank, host wall clock around it.  Equally sized frames go through
    every kernel together (jpgenc_encode_frames / jpgenc_encode_frames_device on one context); `workers` is reported for
    the host threads that build the Huffman tables."""
    from jpgenc_b200.capi import Encoder
    own = enc is None
    if own:
        enc = Encoder(local)
    try:
        # warm-up with the full batch: the context sizes its buffers for the largest pass it has seen
        enc.encode_frames_device(frames_ptrs, w, h, out_ptrs, caps, host_frames=not device_frames)
        t = time.perf_counter()
        for _ in range(reps):
            sizes = enc.encode_frames_device(frames_ptrs, w, h, out_ptrs, caps, host_frames=not device_frames)
        dt = (time.perf_counter() - t) / reps
    finally:
        if own:
            enc.close()
    return len(frames_ptrs) / dt, dt, sizes


def run_batch(args):
    """BASELINE config 4: 1024 frames of 1920x1080, frame k (seed k) encoded by rank owner_of(k); no collective on the
    data path.  value = frames resident in HBM; e2e = frames in pinned host memory, files back in pinned host memory."""
    import torch
    from jpgenc_b200.capi import Encoder, pinned_empty, pinned_free
    from jpgenc_b200.sharding import frames_for_rank

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w, h, desc = WORKLOADS["batch1080p"]
    mine = frames_for_rank(BATCH_FRAMES, rank, world)
    nf, fbytes = len(mine), w * h * 3
    enc = Encoder(local)
    d_all = enc.dev_alloc(nf * fbytes)
    for i, k in enumerate(mine):
        enc.synth_rgb(d_all + i * fbytes, w, h, k)
    enc.synchronize()
    dev_ptrs = [d_all + i * fbytes for i in range(nf)]
    workers = max(2, (os.cpu_count() or 1) // max(1, torch.cuda.device_count()))      # the library's default share of the host per GPU process
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = enc.launch_count()
    reps_dev = max(3, args.steps // 20)
    fps_dev, dt_dev, sizes = batch_frames_per_s(local, dev_ptrs, w, h, workers, True, reps=reps_dev, enc=enc)
    launches = enc.launch_count() - launches0
    barrier()
    clocks = sampler.summary()
    dt_dev = max_over_ranks(dt_dev)
    # end to end: the same frames in pinned host memory, complete files written to pinned host memory
    host, host_ptr = pinned_empty(nf * fbytes)
    enc.d2h(host, d_all)
    cap = max(sizes) + 4096
    out, out_ptr = pinned_empty(nf * cap)
    barrier()
    fps_e2e, dt_e2e, sizes2 = batch_frames_per_s(local, [host_ptr + i * fbytes for i in range(nf)], w, h, workers, False,
                                                 [out_ptr + i * cap for i in range(nf)], [cap] * nf, enc=enc)
    barrier()
    dt_e2e = max_over_ranks(dt_e2e)
    assert sizes2 == sizes and out[0] == 0xFF and out[1] == 0xD8
    mpx = BATCH_FRAMES * w * h / 1e6
    line = {"metric": METRIC, "value": round(mpx / dt_dev, 1), "unit": UNIT, "n_gpus": world, "steps": reps_dev, "warmup": 1,
            "ms_per_step": round(dt_dev * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": desc, "frames": BATCH_FRAMES, "frames_per_gpu": nf, "width": w, "height": h,
                       "mode": "all frames of a pass through every kernel together, passes on three pipeline lanes (jpgenc_encode_frames[_device])",
                       "host_threads_for_tables": workers,
                       "tables_built_on": ("device" if (os.environ.get("JPGENC_DEVICE_TABLES", "") or ("1" if workers < 8 and nf >= 256 else "0")) != "0" else "host"),
                       "timing": "host wall clock around the synchronous batch call, max over ranks",
                       "l2": "inputs larger than L2 (%.1f GB of frames per GPU)" % (nf * fbytes / 1e9)},
            "e2e": {"value": round(mpx / dt_e2e, 1), "unit": UNIT, "h2d_bytes_per_step": nf * fbytes, "d2h_bytes_per_step": int(sum(sizes)),
                    "ms_per_step": round(dt_e2e * 1e3, 3), "frames_per_s": round(BATCH_FRAMES / dt_e2e, 1)},
            "frames_per_s": round(BATCH_FRAMES / dt_dev, 1), "gpu_launches": int(launches), "gpu_launches_note": "6 kernels per pass (a sixth of the rank's frames each), passes are handed to three pipeline lanes; includes the warm-up call",
            "clocks": clocks,
            "jpeg_bytes_per_frame": int(sum(sizes) / max(nf, 1))}
    pinned_free(host_ptr); pinned_free(out_ptr); enc.dev_free(d_all); enc.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    if rank == 0:
        emit(line)


def extra_workloads(enc, args, peak):
    """the other single-GPU configurations of BASELINE.json, measured the same way (reported, not the headline)"""
    import numpy as np
    from jpgenc_b200.tables import ANNEX_K_LUMA
    out = {}
    # configs[1]: DCT + quant + zigzag microbenchmark, 2^24 blocks, 384 algorithmic bytes per block
    nb = 1 << 24
    d_in, d_out = enc.dev_alloc(nb * 256), enc.dev_alloc(nb * 128)
    enc.synth_blocks(d_in, nb)
    qy = ANNEX_K_LUMA
    for _ in range(3):
        refined = enc.dct_quant_blocks(d_in, d_out, nb, qy)
    enc.synchronize()
    enc.timer_begin()
    reps = 10
    for _ in range(reps):
        enc.dct_quant_blocks(d_in, d_out, nb, qy, want_refined=False)
    ms = enc.timer_end() / reps
    out["dct_microbench"] = {"blocks": nb, "ms": round(ms, 4), "gblocks_per_s": round(nb / ms / 1e6, 3),
                             "roofline": {"bound": "hbm", "achieved": round(nb * 384 / (ms / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                                          "frac": round(nb * 384 / (ms / 1e3) / 1e9 / peak, 4)},
                             "refined_blocks": int(refined), "includes": "fast kernel + exact refinement kernel"}
    enc.dev_free(d_in)
    enc.dev_free(d_out)
    # configs[2]: one 3840x2160 frame
    if args.workload != "frame4k":
        w, h = 3840, 2160
        d = enc.dev_alloc(w * h * 3)
        enc.synth_rgb(d, w, h, 0)
        enc.bind_device_rgb(d, w, h)
        for _ in range(3):
            enc.encode_bound(None)
        enc.flush_l2()
        enc.synchronize()
        reps = 20
        tot = 0.0
        for _ in range(reps):
            enc.flush_l2()
            enc.timer_begin()
            enc.encode_bound(None)
            tot += enc.timer_end()
        s = enc.stats()
        out["frame4k"] = {"ms_per_frame": round(tot / reps, 4), "mpx_per_s": round(w * h / 1e6 / (tot / reps / 1e3), 1),
                          "k1_ms": round(s.ms_k1, 4), "l2": "flushed between iterations (256 MB write)"}
        enc.dev_free(d)
    # configs[4] in small: 256 frames of 1920x1080 through the batched-frame calls
    try:
        from jpgenc_b200.capi import pinned_empty, pinned_free
        w, h, nf = 1920, 1080, 256
        fbytes = w * h * 3
        d_all = enc.dev_alloc(nf * fbytes)
        for k in range(nf):
            enc.synth_rgb(d_all + k * fbytes, w, h, k)
        enc.synchronize()
        fps_dev, _, sizes = batch_frames_per_s(enc.device, [d_all + k * fbytes for k in range(nf)], w, h, 8, True, reps=2, enc=enc)
        host, host_ptr = pinned_empty(nf * fbytes)
        enc.d2h(host, d_all)
        cap = max(sizes) + 4096
        outb, out_ptr = pinned_empty(nf * cap)
        fps_e2e, _, _ = batch_frames_per_s(enc.device, [host_ptr + k * fbytes for k in range(nf)], w, h, 8, False,
                                           [out_ptr + k * cap for k in range(nf)], [cap] * nf, reps=2, enc=enc)
        pinned_free(host_ptr); pinned_free(out_ptr); enc.dev_free(d_all)
        out["batch1080p_256"] = {"frames": nf, "mode": "frames batched through the kernels in passes on pipeline lanes (jpgenc_encode_frames[_device])", "device_resident_frames_per_s": round(fps_dev, 1),
                                 "device_resident_mpx_per_s": round(fps_dev * w * h / 1e6, 1),
                                 "e2e_frames_per_s": round(fps_e2e, 1), "e2e_mpx_per_s": round(fps_e2e * w * h / 1e6, 1),
                                 "timing": "host wall clock around the synchronous batch call"}
    except Exception as ex:
        out["batch1080p_256"] = {"error": repr(ex)}
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (NCCL prints its version to stdout at init) must not mix with the one JSON line: everything that
    writes to fd 1 from here on goes to stderr, emit() writes the result to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="image16k", choices=list(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "batch1080p":
        run_batch(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
