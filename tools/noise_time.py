import sys, numpy as np
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
from jpgenc_b200.synth import noise_rgb
E = Encoder(0)
E.set_stage_timing(2)
for n in (4096,):
    rgb = noise_rgb(n, n, 3)
    E.upload_rgb(rgb)
    for i in range(5):
        nb = E.encode_bound(None)
    s = E.stats()
    print(f"noise {n}x{n}: jpeg {nb} B ({8*nb/(n*n):.2f} bit/px) K1 {s.ms_k1:.4f} K1+refine {s.ms_forward:.4f} K2 {s.ms_stats:.4f} K3+K4 {s.ms_entropy:.4f} refined {s.refined_blocks}/{s.n_blocks}")
