"""Pins the oracle against the reference ITSELF (oracle/_ref, compiled by oracle/build_ref.sh from /root/reference).
Runs in the build container; skipped where oracle/_ref was never built.  Not a GPU test."""
import glob
import hashlib
import os

import numpy as np
import pytest

from jpgenc_b200.synth import noise_rgb, ppm_p6_bytes, synth_rgb

FIXTURES = sorted(glob.glob("/root/reference/src/test/res/*.ppm"))


def _ref_jpeg(reference, ppm_bytes, tmp_path, name="in"):
    p, j = tmp_path / f"{name}.ppm", tmp_path / f"{name}.jpg"
    p.write_bytes(ppm_bytes)
    rc, _, _ = reference.encode_file(str(p), str(j))
    assert rc == 0
    return j.read_bytes(), str(p)


@pytest.mark.skipif(not FIXTURES, reason="/root/reference not mounted")
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_reference_fixtures_byte_identical(oracle, reference, tmp_path, path):
    data = open(path, "rb").read()
    ref, _ = _ref_jpeg(reference, data, tmp_path)
    assert oracle.encode_ppm(data) == ref


# SURVEY.md 8(c) pins (generator of 8(d), seed 0)
PINS = {
    (512, 512): ("82eac3f4d5808b2ba53ca02b76058b00619eea319ce5482c933a5f560b16e554", 12323,
                 "ceedd1ecae8dbeaf3620f9256abf57c9061167b65f5d574426739cacd82a3c20"),
    (1920, 1080): ("580cec4112d86c1d5cbf13d6ca129bc01f6afee2026e46d070f480bec135e884", 87008,
                   "f6d1960662d092cef9ff642694f089a6142cb3d3e0d45113e0ec0606c6656554"),
}


@pytest.mark.parametrize("size", list(PINS))
def test_synthetic_pins(oracle, reference, tmp_path, size):
    w, h = size
    ppm = ppm_p6_bytes(synth_rgb(w, h, 0))
    psha, nbytes, jsha = PINS[size]
    assert hashlib.sha256(ppm).hexdigest() == psha
    ref, _ = _ref_jpeg(reference, ppm, tmp_path)
    assert len(ref) == nbytes and hashlib.sha256(ref).hexdigest() == jsha
    assert oracle.encode_ppm(ppm) == ref


@pytest.mark.parametrize("w,h,kind", [(256, 256, "noise"), (100, 37, "noise"), (33, 17, "synth"), (1, 1, "noise"),
                                      (16, 16, "synth"), (17, 16, "synth"), (16, 17, "noise")])
def test_odd_sizes_and_noise(oracle, reference, tmp_path, w, h, kind):
    rgb = synth_rgb(w, h, 5) if kind == "synth" else noise_rgb(w, h, 9)
    ppm = ppm_p6_bytes(rgb)
    ref, _ = _ref_jpeg(reference, ppm, tmp_path)
    assert oracle.encode_ppm(ppm) == ref


def test_stage_dumps_match(oracle, reference, tmp_path):
    """every intermediate plane of the reference, bit for bit (doubles compared exactly)"""
    for name, rgb in {"n": noise_rgb(40, 24, 4), "s": synth_rgb(64, 48, 2)}.items():
        _, path = _ref_jpeg(reference, ppm_p6_bytes(rgb), tmp_path, name)
        r = reference.stage_dump(path)
        o = oracle.forward_planes(rgb)
        for k in ("y", "cb", "cr", "dct_y", "dct_cb", "dct_cr", "q_y", "q_cb", "q_cr"):
            assert np.array_equal(r[k], o[k]), k


def test_dct_and_quantize_blocks(oracle, reference):
    rng = np.random.default_rng(3)
    for _ in range(200):
        blk = rng.uniform(-128, 128, (8, 8))
        assert np.array_equal(oracle.dct(blk), reference.dct(blk))                        # Arai: exact
        for mode in ("direct", "matrix"):
            assert np.allclose(oracle.dct(blk, mode), reference.dct(blk, mode), atol=1e-9)
        d = reference.dct(blk)
        assert np.array_equal(oracle.quantize(d, oracle.qy), reference.quantize(d, oracle.qy))


def test_block_symbols(oracle, reference):
    rng = np.random.default_rng(5)
    for density in (0.02, 0.2, 0.9):
        for _ in range(100):
            blk = (rng.integers(-300, 300, 64) * (rng.random(64) < density)).astype(np.int32)
            a, b = oracle.block_symbols(blk), reference.block_symbols(blk)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_huffman_tables_random_texts(oracle, reference):
    """code lengths, codes AND the order of symbols inside a length (libstdc++ hash/heap order)"""
    rng = np.random.default_rng(11)
    for trial in range(300):
        nsym = int(rng.integers(1, 200))
        alphabet = rng.permutation(256)[:nsym]
        n = int(rng.integers(nsym, 4000))
        if trial % 3 == 0:      # many ties
            text = alphabet[rng.integers(0, nsym, n)]
        else:                   # skewed
            p = rng.random(nsym) ** 4
            text = rng.choice(alphabet, n, p=p / p.sum())
        t = oracle.huffman_from_text(text).as_dict()
        r = reference.huffman(text)
        for k in ("length", "code_msb", "counts", "symbols"):
            assert np.array_equal(t[k], r[k]), (trial, k)


def test_bitstream_pack(oracle, reference):
    rng = np.random.default_rng(13)
    for _ in range(50):
        n = int(rng.integers(1, 300))
        nbits = rng.integers(1, 17, n)
        vals = np.array([int(rng.integers(0, 1 << b)) for b in nbits], np.uint32)
        vals[rng.random(n) < 0.3] = 0xFFFF                   # provoke FF bytes
        vals &= (1 << nbits.astype(np.uint32)) - 1
        msb = vals << (32 - nbits).astype(np.uint32)
        for fill in (False, True):
            a, na = oracle.pack_bits(msb, nbits, msb_aligned=True, fill=fill)
            b, nb = reference.pack_bits(msb, nbits, fill=fill)
            assert na == nb and np.array_equal(a, b)


def test_planes_path_matches_reference(oracle, reference, tmp_path):
    """Image::writeJPEG on an Image assembled from planes of doubles (src/Image.cpp:831-846: RGB planes are converted, an
    image that already is YCbCr is not): jo_encode_planes == the compiled reference, byte for byte"""
    rng = np.random.default_rng(99)
    for (w, h) in [(16, 16), (40, 24), (19, 50)]:
        w16, h16 = (w + 15) // 16 * 16, (h + 15) // 16 * 16
        for ycc in (False, True):
            lo, hi = (-128, 127) if ycc else (0, 255)
            planes = [rng.uniform(lo, hi, (h16, w16)) for _ in range(3)]
            jpg = tmp_path / "p.jpg"
            assert reference.encode_planes(planes[0], planes[1], planes[2], w, h, ycc, str(jpg)) == 0
            assert oracle.encode_planes(planes[0], planes[1], planes[2], w, h, ycc) == jpg.read_bytes(), (w, h, ycc)
    # planes that hold 8-bit samples are the ordinary path
    from jpgenc_b200.synth import synth_rgb
    rgb = synth_rgb(48, 32, 3)
    planes = [rgb[..., k].astype(np.float64) for k in range(3)]
    assert oracle.encode_planes(planes[0], planes[1], planes[2], 48, 32, False) == oracle.encode_rgb(rgb)


def test_stage_methods_on_planes_match_reference(oracle, reference):
    """Image::applySubsampling for every SubsamplingMode (src/Image.cpp:198-319) and Image::applyDCT for every DCTMode
    (src/Image.cpp:540-595, Dct.hpp:47-276) on whole planes: doubles compared exactly"""
    rng = np.random.default_rng(123)
    for (w, h) in [(16, 16), (32, 48), (64, 16)]:
        plane = rng.uniform(-128, 127, (h, w))
        for name, mode in oracle.SUBSAMPLING.items():
            assert np.array_equal(oracle.subsample_plane(plane, name), reference.subsample_plane(plane, mode)), (name, w, h)
        for name, mode in oracle.DCT_MODES.items():
            assert np.array_equal(oracle.dct_plane(plane, name), reference.dct_plane(plane, mode)), (name, w, h)
