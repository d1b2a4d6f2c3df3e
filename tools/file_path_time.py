"""wall time of jpgenc_encode_ppm_file on a 16384^2 P6 file in /dev/shm (page cache) -- development aid"""
import sys, time, hashlib
sys.path.insert(0, ".")
import numpy as np
from jpgenc_b200.capi import Encoder
w = h = 16384
enc = Encoder(0)
d = enc.dev_alloc(w * h * 3)
enc.synth_rgb(d, w, h, 0)
rgb = np.empty(w * h * 3, np.uint8)
enc.d2h(rgb, d)
enc.dev_free(d)
with open("/dev/shm/big.ppm", "wb") as f:
    f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(rgb.tobytes())
for i in range(4):
    t = time.perf_counter()
    enc.encode_ppm_file("/dev/shm/big.ppm", "/dev/shm/big.jpg")
    print(f"encode_ppm_file: {(time.perf_counter()-t)*1e3:.1f} ms")
print(hashlib.sha256(open("/dev/shm/big.jpg","rb").read()).hexdigest())
import os; os.remove("/dev/shm/big.ppm"); os.remove("/dev/shm/big.jpg")
