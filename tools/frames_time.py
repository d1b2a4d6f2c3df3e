"""frames/s of jpgenc_encode_frames_device (all frames through every kernel together) -- development aid"""
import sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder, pinned_empty

w, h = 1920, 1080
fb = w * h * 3
enc = Encoder(0)
enc.set_stage_timing(2)
for nf in (64, 256, 1024):
    d = enc.dev_alloc(nf * fb)
    for k in range(nf):
        enc.synth_rgb(d + k * fb, w, h, k)
    enc.synchronize()
    ptrs = [d + k * fb for k in range(nf)]
    sizes = enc.encode_frames_device(ptrs, w, h)
    t = time.perf_counter(); reps = 3
    for _ in range(reps):
        enc.encode_frames_device(ptrs, w, h)
    dt = (time.perf_counter() - t) / reps
    cap = max(sizes) + 64
    out, op = pinned_empty(nf * cap)
    enc.encode_frames_device(ptrs, w, h, [op + k * cap for k in range(nf)], [cap] * nf)
    t = time.perf_counter()
    for _ in range(reps):
        enc.encode_frames_device(ptrs, w, h, [op + k * cap for k in range(nf)], [cap] * nf)
    dt2 = (time.perf_counter() - t) / reps
    host, hp = pinned_empty(nf * fb)
    enc.d2h(host, d)
    hptrs = [hp + k * fb for k in range(nf)]
    enc.encode_frames_device(hptrs, w, h, [op + k * cap for k in range(nf)], [cap] * nf, host_frames=True)
    t = time.perf_counter()
    for _ in range(reps):
        enc.encode_frames_device(hptrs, w, h, [op + k * cap for k in range(nf)], [cap] * nf, host_frames=True)
    dt3 = (time.perf_counter() - t) / reps
    print(f"   from pinned host memory: {nf/dt3:8.0f} fps = {nf*fb/dt3/1e9:.1f} GB/s H2D")
    s = enc.stats()
    print(f"{nf} frames: sizes only {nf/dt:8.0f} fps ({nf*w*h/1e6/dt:8.0f} Mpx/s), files to pinned host {nf/dt2:8.0f} fps; last pass K1+refine? K2 {s.ms_stats:.3f} K3+K4 {s.ms_entropy:.3f} ms")
    enc.dev_free(d)
