"""The oracle against the known-answer vectors of the reference's own unit tests.

Numbers below are the expected values written in /root/reference/src/test/*.cpp (file:line cited per test);
they are facts about the algorithm, restated here so the check travels without the reference tree.
"""
import numpy as np
import pytest

RAMP = np.arange(1, 65, dtype=np.float64).reshape(8, 8)
# DctTest.cpp:22-31
RAMP_DCT = np.zeros((8, 8))
RAMP_DCT[0, :] = [260, -18.2216411837961, 7.69085915161152e-15, -1.90481782616726, 0, -0.568239222367164,
                  1.85673764701218e-14, -0.143407824981022]
RAMP_DCT[1:, 0] = [-145.773129470369, 0, -15.2385426093380, 0, -4.54591377893732, 0, -1.14726259984816]


@pytest.mark.parametrize("mode", ["arai", "direct", "matrix"])
def test_dct_ramp(oracle, mode):            # DctTest.cpp:9-48, delta 1e-5 (unittest.hpp:12-24)
    assert np.allclose(oracle.dct(RAMP, mode), RAMP_DCT, atol=1e-5, rtol=0)


def test_zigzag_permutation(oracle):        # DctTest.cpp:86-109
    exp = [1, 2, 9, 17, 10, 3, 4, 11, 18, 25, 33, 26, 19, 12, 5, 6, 13, 20, 27, 34, 41, 49, 42, 35, 28, 21, 14, 7, 8, 15,
           22, 29, 36, 43, 50, 57, 58, 51, 44, 37, 30, 23, 16, 24, 31, 38, 45, 52, 59, 60, 53, 46, 39, 32, 40, 47, 54, 61,
           62, 55, 48, 56, 63, 64]
    flat = RAMP.reshape(64)
    assert [int(flat[oracle.zigzag_index(i)]) for i in range(64)] == exp


def test_quantization(oracle):              # DctTest.cpp:112-158
    inp = [581, -144, 56, 17, 15, -7, 25, -9, -242, 133, -48, 42, -2, -7, 13, -4, 108, -18, -40, 71, -33, 12, 6, -10,
           -56, -93, 48, 19, -8, 7, 6, -2, -17, 9, 7, -23, -3, -10, 5, 3, 4, 9, -4, -5, 2, 2, -7, 3, -9, 7, 8, -6, 5, 12,
           2, -5, -9, -4, -2, -3, 6, 1, -1, -1]
    exp = [36, -13, 6, 1, 1, 0, 0, 0, -20, 11, -3, 2, 0, 0, 0, 0, 8, -1, -3, 3, -1, 0, 0, 0, -4, -5, 2, 1, 0, 0, 0, 0,
           -1, 0, 0, 0, 0, 0, 0, 0] + [0] * 24
    assert oracle.quantize(np.array(inp, float), oracle.qy).reshape(64).tolist() == exp


def test_quantize_rounds_half_away_from_zero(oracle):   # Coding.hpp:93 std::round
    blk = np.zeros(64)
    blk[0], blk[1], blk[2], blk[3] = 8.0, -8.0, 24.0, -5.5      # /16 -> .5 ; /11 ; /10 -> 2.4 ; /16
    q = oracle.quantize(blk, oracle.qy).reshape(64)
    assert q[0] == 1 and q[1] == -1 and q[2] == 2 and q[3] == 0


NATURAL = np.zeros(64, np.int32)
NATURAL[[0, 1, 20, 25, 63]] = [-111, 57, 3, -2, -2]


def test_rle_with_zigzag(oracle):           # CodingTest.cpp:70-131: zigzag gives -111 57 9*0 -2 13*0 3 37*0 -2
    sym, bits, nb = oracle.block_symbols(NATURAL)
    #  (0,-111) (0,57) (9,-2) (13,3) ZRL ZRL (5,-2)
    assert sym.tolist() == [7, 6, (9 << 4) | 2, (13 << 4) | 2, 0xF0, 0xF0, (5 << 4) | 2]
    assert nb.tolist() == [7, 6, 2, 2, 0, 0, 2]
    assert bits.tolist() == [16, 57, 1, 3, 0, 0, 1]       # -111 -> 127-111, -2 -> 3-2 (CodingTest.cpp:57-64)
    z = NATURAL.copy(); z[63] = 0                           # trailing zeros -> EOB
    sym, bits, nb = oracle.block_symbols(z)
    assert sym.tolist() == [7, 6, (9 << 4) | 2, (13 << 4) | 2, 0x00]


def test_rle_symbols_linear_order(oracle):  # CodingTest.cpp:5-68 expects 7,6,240,34,66,240,240,82 on the LINEAR list
    # build a natural-order block whose zigzag sequence equals that linear list: -111 57 18*0 3 4*0 -2 37*0 -2
    seq = np.zeros(64, np.int32)
    seq[[0, 1, 20, 25, 63]] = [-111, 57, 3, -2, -2]
    nat = np.zeros(64, np.int32)
    for i in range(64):
        nat[oracle.zigzag_index(i)] = seq[i]
    sym, bits, nb = oracle.block_symbols(nat)
    assert sym.tolist() == [7, 6, 240, 34, 66, 240, 240, 82]
    assert list(zip(bits.tolist(), nb.tolist())) == [(16, 7), (57, 6), (0, 0), (3, 2), (1, 2), (0, 0), (0, 0), (1, 2)]


@pytest.mark.parametrize("value,cat,bits", [            # CodingTest.cpp:133-162
    (0, 0, 0), (-1, 1, 0), (1, 1, 1), (-3, 2, 0), (-2, 2, 1), (2, 2, 2), (3, 2, 3), (-7, 3, 0), (-6, 3, 1), (-4, 3, 3),
    (4, 3, 4), (6, 3, 6), (7, 3, 7), (-1023, 10, 0), (-1022, 10, 1), (-512, 10, 511), (512, 10, 512), (1022, 10, 1022),
    (1023, 10, 1023)])
def test_category(oracle, value, cat, bits):
    assert oracle.category(value) == (cat, bits)


def test_colour_values(oracle):             # ImageTest.cpp:47-57: (255,0,255) -> Y -22.685 Cb 84.4815 Cr 106.7685
    rgb = np.zeros((16, 16, 3), np.uint8)
    rgb[...] = (255, 0, 255)
    d = oracle.forward_planes(rgb)
    assert abs(d["y"][0, 0] - (-22.685)) < 1e-5
    assert abs(d["cb"][0, 0] - 84.4815) < 1e-5
    assert abs(d["cr"][0, 0] - 106.7685) < 1e-5


def test_s420m_is_mean_of_four(oracle):     # ImageTest.cpp:154-175 (means of 2x2 neighbourhoods)
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    d = oracle.forward_planes(rgb)
    r, g, b = (rgb[..., i].astype(np.float64) for i in range(3))
    cb = -np.float32(.1687) * r - np.float32(.3312) * g + np.float32(.5) * b
    mean = (cb[0::2, 0::2] + cb[0::2, 1::2] + cb[1::2, 0::2] + cb[1::2, 1::2]) / 4
    assert np.allclose(d["cb"], mean, atol=1e-9)


def test_padding_replicates_edges(oracle):  # ImageTest.cpp:25-44: (15,0),(0,15),(15,15) copy the border
    rgb = np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 4, 3) * 5
    d = oracle.forward_planes(rgb)
    y = d["y"]
    assert y.shape == (16, 16)
    assert y[0, 15] == y[0, 3] and y[15, 0] == y[3, 0] and y[15, 15] == y[3, 3]


def test_p3_maxval_scaling(oracle):         # ImageTest.cpp:7-24: P3 maxval 15 -> x 255/15
    ppm = b"P3\n# c\n2 1\n15\n15 0 3  0 15 15\n"
    rgb, maxval = oracle.ppm_load(ppm)
    assert maxval == 15 and rgb.reshape(-1).tolist() == [15, 0, 3, 0, 15, 15]
    d = oracle.forward_planes(rgb, maxval)
    r, g, b = 255.0, 0.0, 3 * (255. / 15)
    f = lambda c: float(np.float32(c))      # float constants widened to double, Image.cpp:131-134
    assert abs(d["y"][0, 0] - (f(.299) * r + f(.587) * g + f(.114) * b - 128)) < 1e-9


def test_bitstream_append_and_fill(oracle):  # BitstreamGenericTest.cpp:49-68,108-127
    out, nbits = oracle.pack_bits([0x34000000], [6], msb_aligned=True, fill=False)   # push_back(0x34000000, 6) -> 001101
    assert nbits == 6 and out.tolist() == [0b00110100]
    out, nbits = oracle.pack_bits([0x34000000], [6], msb_aligned=True, fill=True)
    assert nbits == 8 and out.tolist() == [0b00110111]
    out, nbits = oracle.pack_bits([0xB, 0x0C0], [4, 12], fill=False)                 # concat -> 0xB0C0
    assert out.tolist() == [0xB0, 0xC0]
    out, nbits = oracle.pack_bits([0xFF, 0x1], [8, 1], fill=True)                    # FF is stuffed, pad byte FF too
    assert out.tolist() == [0xFF, 0x00, 0xFF, 0x00]


def test_header_layout(oracle):             # ImageTest.cpp:302,319: sizeof(sAPP0)==18, sizeof(sSOF0)==19
    tabs, _, _, _ = oracle.entropy_encode(np.zeros((1, 6, 64), np.int16), 1, 1)
    h = oracle.headers(300, 200, tabs).tobytes()
    assert h[:2] == b"\xff\xd8" and h[2:4] == b"\xff\xe0" and h[4:6] == b"\x00\x10" and h[6:11] == b"JFIF\0"
    app0_end = 2 + 18
    assert h[app0_end:app0_end + 2] == b"\xff\xdb"
    sof = h.index(b"\xff\xc0")
    assert h[sof + 2:sof + 4] == b"\x00\x11" and h[sof + 4] == 8
    assert h[sof + 5:sof + 9] == bytes([0, 200, 1, 44])           # Y then X, hi/lo
    assert h[sof + 10:sof + 19] == bytes([1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1])
    assert h[-14:] == bytes([0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3F, 0])
