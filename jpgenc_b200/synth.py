"""Synthetic inputs for the jpgEnc hot path.

Integer-only generators so that the CPU (numpy) and any other implementation agree bit for bit.
`synth_rgb(w, h, seed)` is the gradient + checker + hash-noise image the BASELINE pins were made with;
`noise_rgb(w, h, seed)` is the high-entropy variant (uniform u8) that stresses Huffman pack / stuffing.
"""
from __future__ import annotations

import numpy as np

__all__ = ["synth_rgb", "noise_rgb", "ppm_p6_bytes", "ppm_p3_bytes", "write_ppm"]


def _hash32(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & 0xFFFFFFFF
    v ^= v >> 16
    v = (v * 0x7FEB352D) & 0xFFFFFFFF
    v ^= v >> 15
    v = (v * 0x846CA68B) & 0xFFFFFFFF
    v ^= v >> 16
    return v


def synth_rgb(w: int, h: int, seed: int = 0, rows: slice | None = None) -> np.ndarray:
    """(h, w, 3) uint8.  `rows` restricts generation to a row band (for huge images)."""
    r0, r1 = (0, h) if rows is None else (rows.start or 0, rows.stop if rows.stop is not None else h)
    y = np.arange(r0, r1, dtype=np.uint64)[:, None, None]
    x = np.arange(w, dtype=np.uint64)[None, :, None]
    c = np.arange(3, dtype=np.uint64)[None, None, :]
    v = ((y * w + x) * 3 + c + (seed * 0x9E3779B9)) & 0xFFFFFFFF
    noise = (_hash32(v) & 31).astype(np.int64) - 16
    xi = x.astype(np.int64)
    yi = y.astype(np.int64)
    base = np.empty((r1 - r0, w, 3), dtype=np.int64)
    base[..., 0] = (255 * xi[..., 0]) // max(w - 1, 1)
    base[..., 1] = (255 * yi[..., 0]) // max(h - 1, 1)
    base[..., 2] = (255 * (xi[..., 0] + yi[..., 0])) // max(w + h - 2, 1)
    chk = 40 * (((xi >> 5) ^ (yi >> 5)) & 1)
    return np.clip(base + chk + noise, 0, 255).astype(np.uint8)


def noise_rgb(w: int, h: int, seed: int = 1) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def ppm_p6_bytes(rgb: np.ndarray, maxval: int = 255) -> bytes:
    h, w, _ = rgb.shape
    return b"P6\n%d %d\n%d\n" % (w, h, maxval) + np.ascontiguousarray(rgb, dtype=np.uint8).tobytes()


def ppm_p3_bytes(rgb: np.ndarray, maxval: int = 255, comment: str | None = None) -> bytes:
    h, w, _ = rgb.shape
    head = "P3\n" + (f"# {comment}\n" if comment else "") + f"{w} {h}\n{maxval}\n"
    body = "\n".join(" ".join(str(int(v)) for v in row.reshape(-1)) for row in rgb) + "\n"
    return (head + body).encode()


def write_ppm(path: str, rgb: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(ppm_p6_bytes(rgb))
