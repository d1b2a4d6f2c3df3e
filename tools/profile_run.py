"""Short fixed workload for ncu: a few whole encodes of one synthetic image (and optionally the DCT microbenchmark)."""
import argparse
import sys

sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
from jpgenc_b200.tables import ANNEX_K_LUMA

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=16384)
ap.add_argument("--h", type=int, default=16384)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--dct-blocks", type=int, default=0)
a = ap.parse_args()
E = Encoder(0)
E.set_stage_timing(2)
d = E.dev_alloc(a.w * a.h * 3)
E.synth_rgb(d, a.w, a.h, 0)
E.bind_device_rgb(d, a.w, a.h)
for i in range(a.iters):
    n = E.encode_bound(None)
s = E.stats()
print(f"{a.w}x{a.h}: jpeg {n} B  K1 {s.ms_k1:.4f} K1+refine {s.ms_forward:.4f} K2 {s.ms_stats:.4f} K3+K4 {s.ms_entropy:.4f} refined {s.refined_blocks}")
if a.dct_blocks:
    di, do = E.dev_alloc(a.dct_blocks * 256), E.dev_alloc(a.dct_blocks * 128)
    E.synth_blocks(di, a.dct_blocks)
    for i in range(a.iters):
        r = E.dct_quant_blocks(di, do, a.dct_blocks, ANNEX_K_LUMA)
    print("dct blocks", a.dct_blocks, "refined", r)
