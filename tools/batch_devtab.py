"""batched frames with the tables built on the host vs on the device -- development aid"""
import os, sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
w, h, nf = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
for dev in ("0", "1"):
    os.environ["JPGENC_DEVICE_TABLES"] = dev
    for lanes in ("3", "4"):
        os.environ["JPGENC_LANES"] = lanes
        ref = enc.encode_frames_device(ptrs, w, h)
        t = time.perf_counter()
        for _ in range(5):
            enc.encode_frames_device(ptrs, w, h)
        dt = (time.perf_counter() - t) / 5
        print(f"device tables {dev} lanes {lanes}: {nf/dt:8.0f} fps", flush=True)
os.environ["JPGENC_TRACE"] = "1"
os.environ["JPGENC_LANES"] = "1"; os.environ["JPGENC_FRAMES_PER_PASS"] = "103"
for dev in ("0", "1"):
    os.environ["JPGENC_DEVICE_TABLES"] = dev
    enc.encode_frames_device(ptrs[:min(nf, 206)], w, h)
    print("--- trace, one lane, device tables", dev, file=sys.stderr, flush=True)
    enc.encode_frames_device(ptrs[:min(nf, 206)], w, h)
