"""frames/s of the batched-frame call against pass size and slot count (JPGENC_FRAMES_PER_PASS / JPGENC_SLOTS) -- development aid
    python tools/batch_sweep.py [frames] [slots list] [per-pass list]"""
import os, sys, time
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder, pinned_empty

w, h, nf = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
slots_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4, 6]
per_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 32, 64, 128]
fb = w * h * 3
enc = Encoder(0)
d = enc.dev_alloc(nf * fb)
for k in range(nf):
    enc.synth_rgb(d + k * fb, w, h, k)
enc.synchronize()
ptrs = [d + k * fb for k in range(nf)]
host, hp = pinned_empty(nf * fb)
enc.d2h(host, d)
hptrs = [hp + k * fb for k in range(nf)]
offs, sizes, total = enc.encode_frames_packed(ptrs, w, h, None, 0)
out, op = pinned_empty(total + 4096)


def rate(fn, reps):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return nf / ((time.perf_counter() - t) / reps)


for slots in slots_list:
    for per in per_list:
        os.environ["JPGENC_SLOTS"] = str(slots)
        os.environ["JPGENC_FRAMES_PER_PASS"] = str(per)
        reps = 5 if nf >= 512 else 20
        r1 = rate(lambda: enc.encode_frames_packed(ptrs, w, h, None, 0), reps)
        r2 = rate(lambda: enc.encode_frames_packed(ptrs, w, h, op, out.size), reps)
        line = f"frames {nf} slots {slots} per_pass {per:4d}: resident {r1:8.0f} fps, files to pinned host {r2:8.0f} fps"
        if slots == slots_list[-1] or per == 32:
            r3 = rate(lambda: enc.encode_frames_packed(hptrs, w, h, op, out.size, host_frames=True), 2)
            line += f", from pinned host {r3:8.0f} fps = {r3*fb/1e9:.1f} GB/s"
        print(line, flush=True)
