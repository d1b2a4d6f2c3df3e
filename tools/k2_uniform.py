"""K2 on a uniform texture (every block identical: every lane of a warp meets the same symbols) -- development aid"""
import sys
sys.path.insert(0, ".")
import numpy as np
from jpgenc_b200.capi import Encoder
enc = Encoder(0)
enc.set_stage_timing(2)
w = h = 8192
yy, xx = np.mgrid[0:16, 0:16]
cell = np.stack([(xx * 16) % 256, (yy * 9) % 256, ((xx + yy) * 7) % 256], -1).astype(np.uint8)
rgb = np.tile(cell, (h // 16, w // 16, 1))
d = enc.dev_alloc(w * h * 3)
enc.h2d(d, rgb)
enc.bind_device_rgb(d, w, h)
for _ in range(3):
    n = enc.encode_bound(None)
enc.synchronize()
enc.timer_begin()
for _ in range(20):
    enc.encode_bound(None)
ms = enc.timer_end() / 20
s = enc.stats()
print(f"uniform {w}x{h}: {ms:.4f} ms, {n} bytes; k2 {s.ms_stats:.4f} k3k4 {s.ms_entropy:.4f}")
