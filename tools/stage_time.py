"""stage times on the 16384^2 synthetic and on a 4K frame -- development aid"""
import sys
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
enc = Encoder(0)
enc.set_stage_timing(2)
for w, h in ((16384, 16384), (3840, 2160)):
    d = enc.dev_alloc(w * h * 3)
    enc.synth_rgb(d, w, h, 0)
    enc.bind_device_rgb(d, w, h)
    for _ in range(3):
        n = enc.encode_bound(None)
    k = {"k1": [], "fwd": [], "k2": [], "ent": [], "tot": []}
    for _ in range(30):
        enc.timer_begin()
        enc.encode_bound(None)
        k["tot"].append(enc.timer_end())
        s = enc.stats()
        k["k1"].append(s.ms_k1); k["fwd"].append(s.ms_forward); k["k2"].append(s.ms_stats); k["ent"].append(s.ms_entropy)
    med = {a: sorted(v)[15] for a, v in k.items()}
    print(f"{w}x{h}: K1 {med['k1']:.4f} K1+refine {med['fwd']:.4f} K2 {med['k2']:.4f} K3+K4 {med['ent']:.4f} whole {med['tot']:.4f} ms; {n} bytes")
    enc.dev_free(d)
