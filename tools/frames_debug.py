import sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
from bench import ClockSampler
w, h, nf = 1920, 1080, 256
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
def run(tag):
    enc.encode_frames_device(ptrs[:64], w, h)
    t = time.perf_counter()
    for _ in range(2):
        enc.encode_frames_device(ptrs, w, h)
    dt = (time.perf_counter() - t) / 2
    print(f"{tag}: {nf/dt:8.0f} fps")
run("fresh context")
big = enc.dev_alloc(16384 * 16384 * 3)
enc.synth_rgb(big, 16384, 16384, 0)
enc.bind_device_rgb(big, 16384, 16384)
for _ in range(3):
    enc.encode_bound(None)
run("after a 16384^2 image on the same context")
enc.flush_l2(); enc.synchronize()
run("after flush_l2")
s = ClockSampler(0); s.start()
run("with the nvidia-smi sampler thread running")
print(s.summary())
run("sampler stopped")
