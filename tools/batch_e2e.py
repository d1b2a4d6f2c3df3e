"""host-frame batch rate with and without returning the files, contiguous and scattered frames -- development aid"""
import os, sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder, pinned_empty
w, h, nf = 1920, 1080, 1024
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
host, hp = pinned_empty(nf * fb + nf * 4096)
enc.d2h(host[: nf * fb], d)
sizes = enc.encode_frames_device([d + k * fb for k in range(nf)], w, h)
cap = max(sizes) + 64
out, op = pinned_empty(nf * cap)
optrs = [op + k * cap for k in range(nf)]
def run(ptrs, outs, label):
    enc.encode_frames_device(ptrs, w, h, outs, [cap] * nf if outs else None, host_frames=True)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        enc.encode_frames_device(ptrs, w, h, outs, [cap] * nf if outs else None, host_frames=True)
        best = min(best, time.perf_counter() - t)
    print(f"{label}: {nf/best:8.0f} fps = {nf*fb/best/1e9:.2f} GB/s H2D ({best*1e3:.1f} ms)", flush=True)
contig = [hp + k * fb for k in range(nf)]
for per in (os.environ.get("JPGENC_FRAMES_PER_PASS", "") or "41", "16", "96"):
    os.environ["JPGENC_FRAMES_PER_PASS"] = per
    run(contig, optrs, f"per_pass {per} contiguous frames, files returned")
    run(contig, None, f"per_pass {per} contiguous frames, sizes only   ")
# scattered: spread the frames 4 KB apart so that no two are adjacent
import numpy as np
for k in range(nf - 1, -1, -1):
    host[k * (fb + 4096): k * (fb + 4096) + fb] = host[k * fb: (k + 1) * fb]
scat = [hp + k * (fb + 4096) for k in range(nf)]
os.environ["JPGENC_FRAMES_PER_PASS"] = "41"
run(scat, optrs, "per_pass 41 scattered frames, files returned ")
