"""per-pass finish times of the batched-frame call (JPGENC_TRACE=1), sizes only and with the files returned -- development aid"""
import os, sys, time
os.environ["JPGENC_TRACE"] = "1"
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder, pinned_empty
w, h, nf = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
offs, sizes, total = enc.encode_frames_packed(ptrs, w, h, None, 0)
out, op = pinned_empty(total + 4096)
for what, args in (("sizes only", (None, 0)), ("files returned", (op, out.size))):
    enc.encode_frames_packed(ptrs, w, h, *args)
    print(f"--- {what}", file=sys.stderr, flush=True)
    t = time.perf_counter()
    enc.encode_frames_packed(ptrs, w, h, *args)
    print(f"--- total {(time.perf_counter()-t)*1e3:.2f} ms", file=sys.stderr, flush=True)
