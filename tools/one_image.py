"""a few whole encodes of one synthetic image, device resident (development aid for traces / ncu runs)
    python tools/one_image.py W H [reps] [stage timing level 0..2, default 0: none]"""
import sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
w, h = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
level = int(sys.argv[4]) if len(sys.argv) > 4 else 0
enc = Encoder(0)
enc.set_stage_timing(level)
d = enc.dev_alloc(w * h * 3)
enc.synth_rgb(d, w, h, 0)
enc.bind_device_rgb(d, w, h)
for _ in range(3):
    n = enc.encode_bound(None)
enc.synchronize()
enc.timer_begin()
for _ in range(reps):
    n = enc.encode_bound(None)
ms = enc.timer_end() / reps
s = enc.stats()
print(f"{w}x{h}: {ms:.4f} ms per encode, {n} bytes; k1 {s.ms_k1:.4f} fwd {s.ms_forward:.4f} k2 {s.ms_stats:.4f} k3k4 {s.ms_entropy:.4f}")
