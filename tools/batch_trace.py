"""per-pass phase times of the batched-frame call (JPGENC_TRACE=1) -- development aid"""
import os, sys, time
os.environ["JPGENC_TRACE"] = "1"
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
w, h, nf = 1920, 1080, 1024
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
for lanes, per in ((3, 103), (4, 64)):
    os.environ["JPGENC_LANES"] = str(lanes); os.environ["JPGENC_FRAMES_PER_PASS"] = str(per)
    enc.encode_frames_device(ptrs, w, h)
    enc.encode_frames_device(ptrs, w, h)
    print(f"--- lanes {lanes} per_pass {per}", file=sys.stderr, flush=True)
    t = time.perf_counter()
    enc.encode_frames_device(ptrs, w, h)
    print(f"--- total {(time.perf_counter()-t)*1e3:.2f} ms", file=sys.stderr, flush=True)
