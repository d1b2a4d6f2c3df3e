"""pinned host -> device copy rate for one big copy and for band-sized pieces -- development aid"""
import sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder, pinned_empty
enc = Encoder(0)
n = 16384 * 16384 * 3
host, hp = pinned_empty(n)
host[:] = 7
d = enc.dev_alloc(n)
import numpy as np
for pieces in (1, 4, 16, 64, 256):
    step = n // pieces
    best = 1e9
    for _ in range(5):
        enc.synchronize()
        t = time.perf_counter()
        for k in range(pieces):
            enc.lib.jpgenc_memcpy_h2d(enc.h, d + k * step, hp + k * step, step)
        enc.synchronize()
        best = min(best, time.perf_counter() - t)
    print(f"{pieces:4d} pieces: {n/best/1e9:.2f} GB/s ({best*1e3:.3f} ms)")
