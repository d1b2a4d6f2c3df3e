"""frames/s of the batch API for several worker counts (development aid)"""
import sys, time
sys.path.insert(0, ".")
from bench import batch_frames_per_s
from jpgenc_b200.capi import Encoder, pinned_empty

w, h, nf = 1920, 1080, 256
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
host, hp = pinned_empty(nf * fb)
enc.d2h(host, d)
cap = 200000
out, op = pinned_empty(nf * cap)
for workers in (2, 4, 8, 12, 16, 24):
    fd, _, sizes = batch_frames_per_s(0, [d + k * fb for k in range(nf)], w, h, workers, True, reps=2)
    fe, _, _ = batch_frames_per_s(0, [hp + k * fb for k in range(nf)], w, h, workers, False, [op + k * cap for k in range(nf)], [cap] * nf, reps=2)
    print(f"workers {workers:2d}: device-resident {fd:8.0f} fps, e2e {fe:8.0f} fps ({fe*fb/1e9:.1f} GB/s H2D)")
