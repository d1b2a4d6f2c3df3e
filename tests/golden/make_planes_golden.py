"""Generates tests/golden/planes.npz from the COMPILED REFERENCE (oracle/_ref): Image::writeJPEG on Images assembled in memory
from planes of doubles -- edited 8-bit samples, non-integral values, planes that already are YCbCr.  Run in the build
container only; the .npz is committed so that the GPU box can check jpgenc_encode_planes against the reference's own bytes.

    python tests/golden/make_planes_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference  # noqa: E402


def cases():
    rng = np.random.default_rng(4242)
    out = {}
    for name, (w, h), ycc, kind in [("rgb_real_48x32", (48, 32), False, "real"), ("ycc_real_21x35", (21, 35), True, "real"),
                                    ("rgb_edited_64x64", (64, 64), False, "edited"), ("ycc_edited_30x17", (30, 17), True, "edited"),
                                    ("rgb_real_16x16", (16, 16), False, "real")]:
        w16, h16 = (w + 15) // 16 * 16, (h + 15) // 16 * 16
        if kind == "real":
            lo, hi = (-128, 127) if ycc else (0, 255)
            planes = [rng.uniform(lo, hi, (h16, w16)) for _ in range(3)]
        else:
            planes = [rng.integers(0, 256, (h16, w16)).astype(np.float64) - (128 if ycc else 0) for _ in range(3)]
            planes[0][3, 5] += 0.37                      # one edited sample, one edited padding sample
            planes[2][h16 - 1, w16 - 1] = 99.5
        out[name] = (planes, w, h, ycc)
    return out


def main():
    R = Reference()
    tmp = tempfile.mkdtemp()
    out, names = {}, []
    for name, (planes, w, h, ycc) in cases().items():
        j = os.path.join(tmp, name + ".jpg")
        assert R.encode_planes(planes[0], planes[1], planes[2], w, h, ycc, j) == 0
        out[f"{name}/planes"] = np.stack(planes)
        out[f"{name}/dims"] = np.array([w, h, int(ycc)], np.int64)
        out[f"{name}/jpg"] = np.frombuffer(open(j, "rb").read(), np.uint8)
        names.append(name)
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "planes.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
