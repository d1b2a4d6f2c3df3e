"""Generates tests/golden/vectors.npz from the COMPILED REFERENCE (oracle/_ref, built by oracle/build_ref.sh
from /root/reference).  Run in the build container only; the .npz is committed so that the GPU box, which has no
/root/reference, can still check the oracle and the CUDA path against the reference's own outputs.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from jpgenc_b200.synth import noise_rgb, ppm_p3_bytes, ppm_p6_bytes, synth_rgb  # noqa: E402
from oracle.pyoracle import Reference  # noqa: E402


def cases():
    rng = np.random.default_rng(2026)
    text = np.zeros((32, 32, 3), np.uint8)
    text[4:28:3, 2:30] = 255
    text[:, 7::9] = (255, 40, 0)
    grad = np.stack(np.meshgrid(np.arange(26) * 9, np.arange(19) * 13), -1)
    grad = np.concatenate([grad, (grad[..., :1] + grad[..., 1:]) // 2], -1).astype(np.uint8)
    return {
        "synth_64x48": ppm_p6_bytes(synth_rgb(64, 48, 7)),
        "noise_40x24": ppm_p6_bytes(noise_rgb(40, 24, 11)),
        "grad_26x19": ppm_p6_bytes(grad),
        "flat_16x16": ppm_p6_bytes(np.full((16, 16, 3), (200, 30, 90), np.uint8)),
        "tiny_5x3": ppm_p6_bytes(rng.integers(0, 256, (3, 5, 3), dtype=np.uint8)),
        "text_32x32": ppm_p6_bytes(text),
        "p3_12x8_max15": ppm_p3_bytes(rng.integers(0, 16, (8, 12, 3), dtype=np.uint8), 15, comment="made by make_golden.py"),
        "p6_max63": ppm_p6_bytes(rng.integers(0, 64, (17, 33, 3), dtype=np.uint8), 63),
        "synth_128x128": ppm_p6_bytes(synth_rgb(128, 128, 0)),
        "noise_96x80": ppm_p6_bytes(noise_rgb(96, 80, 3)),
    }


def main():
    R = Reference()
    out = {}
    tmp = tempfile.mkdtemp()
    names = []
    for name, ppm in cases().items():
        p = os.path.join(tmp, name + ".ppm")
        j = os.path.join(tmp, name + ".jpg")
        open(p, "wb").write(ppm)
        rc, _, _ = R.encode_file(p, j)
        assert rc == 0
        d = R.stage_dump(p)
        out[f"{name}/ppm"] = np.frombuffer(ppm, np.uint8)
        out[f"{name}/jpg"] = np.frombuffer(open(j, "rb").read(), np.uint8)
        out[f"{name}/q_y"] = d["q_y"].astype(np.int16)
        out[f"{name}/q_cb"] = d["q_cb"].astype(np.int16)
        out[f"{name}/q_cr"] = d["q_cr"].astype(np.int16)
        if d["y"].size <= 64 * 48:
            out[f"{name}/y"] = d["y"]
            out[f"{name}/cb"] = d["cb"]
            out[f"{name}/dct_y"] = d["dct_y"]
        names.append(name)
        print(name, len(ppm), "->", out[f"{name}/jpg"].size, hashlib.sha256(out[f"{name}/jpg"].tobytes()).hexdigest()[:16])
    # Huffman tables of the reference for a few symbol texts (tie-heavy, skewed, single symbol, > 30 symbols)
    rng = np.random.default_rng(7)
    texts = {
        "single": np.full(9, 5),
        "two": np.array([3, 3, 3, 250]),
        "ties": np.repeat(np.arange(12), 4)[rng.permutation(48)],
        "skewed": rng.geometric(0.25, 4000).clip(1, 40),
        "wide": rng.integers(0, 200, 5000),
        "jpeg_like": np.concatenate([np.zeros(3000, int), rng.choice([1, 2, 17, 33, 0xF0, 18, 49], 2000)]),
    }
    for k, t in texts.items():
        h = R.huffman(t)
        out[f"huff/{k}/text"] = t.astype(np.int32)
        out[f"huff/{k}/length"] = h["length"]
        out[f"huff/{k}/code_msb"] = h["code_msb"]
        out[f"huff/{k}/counts"] = h["counts"]
        out[f"huff/{k}/symbols"] = h["symbols"]
    out["names"] = np.array(names)
    out["huff_names"] = np.array(list(texts))
    path = os.path.join(os.path.dirname(__file__), "vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
