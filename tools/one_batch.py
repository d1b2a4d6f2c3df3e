"""one warm-up call + `reps` calls of the batched-frame path on `frames` 1920x1080 frames (development aid for ncu runs)
    python tools/one_batch.py [frames] [reps]"""
import sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder

w, h = 1920, 1080
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
enc.encode_frames_packed(ptrs, w, h, None, 0)
t = time.perf_counter()
for _ in range(reps):
    offs, sizes, total = enc.encode_frames_packed(ptrs, w, h, None, 0)
print(f"{nf} frames: {(time.perf_counter() - t) / reps * 1e3:.3f} ms per call, {total} bytes")
