"""frames/s of the batched-frame calls against pass size and lane count (JPGENC_FRAMES_PER_PASS / JPGENC_LANES) -- development aid"""
import os, sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder, pinned_empty

w, h, nf = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
host, hp = pinned_empty(nf * fb)
enc.d2h(host, d)
hptrs = [hp + k * fb for k in range(nf)]
sizes = enc.encode_frames_device(ptrs, w, h)
cap = max(sizes) + 64
out, op = pinned_empty(nf * cap)
optrs = [op + k * cap for k in range(nf)]
for lanes in (1, 2, 3, 4):
    for per in (171, 128, 103, 64, 48):
        os.environ["JPGENC_LANES"] = str(lanes)
        os.environ["JPGENC_FRAMES_PER_PASS"] = str(per)
        enc.encode_frames_device(ptrs, w, h)
        reps = 5
        t = time.perf_counter()
        for _ in range(reps):
            enc.encode_frames_device(ptrs, w, h)
        dt = (time.perf_counter() - t) / reps
        enc.encode_frames_device(ptrs, w, h, optrs, [cap] * nf)
        t = time.perf_counter()
        for _ in range(reps):
            enc.encode_frames_device(ptrs, w, h, optrs, [cap] * nf)
        dt2 = (time.perf_counter() - t) / reps
        line = f"lanes {lanes} per_pass {per:4d}: resident {nf/dt:8.0f} fps, files to pinned host {nf/dt2:8.0f} fps"
        if per in (341, 128, 41, 32, 64):
            enc.encode_frames_device(hptrs, w, h, optrs, [cap] * nf, host_frames=True)
            t = time.perf_counter()
            for _ in range(2):
                enc.encode_frames_device(hptrs, w, h, optrs, [cap] * nf, host_frames=True)
            dt3 = (time.perf_counter() - t) / 2
            line += f", from pinned host {nf/dt3:8.0f} fps = {nf*fb/dt3/1e9:.1f} GB/s"
        print(line, flush=True)
