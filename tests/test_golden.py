"""The oracle against the committed golden vectors (tests/golden/vectors.npz, produced from the compiled reference
by tests/golden/make_golden.py).  Travels to the GPU box, where /root/reference does not exist."""
import numpy as np
import pytest


def _names(golden, key):
    return [str(n) for n in golden[key]]


def test_golden_files_byte_identical(oracle, golden):
    for name in _names(golden, "names"):
        assert oracle.encode_ppm(golden[f"{name}/ppm"].tobytes()) == golden[f"{name}/jpg"].tobytes(), name


def test_golden_coefficients(oracle, golden):
    for name in _names(golden, "names"):
        rgb, maxval = oracle.ppm_load(golden[f"{name}/ppm"].tobytes())
        d = oracle.forward_planes(rgb, maxval)
        for k in ("q_y", "q_cb", "q_cr"):
            assert np.array_equal(d[k], golden[f"{name}/{k}"].astype(np.int32)), (name, k)
        if f"{name}/y" in golden:
            for k in ("y", "cb", "dct_y"):
                assert np.array_equal(d[k], golden[f"{name}/{k}"]), (name, k)     # doubles, exact
        # MCU-ordered forward output == re-ordered planes
        mcu = oracle.forward(rgb, maxval)
        assert np.array_equal(mcu, oracle.planes_to_mcu(d["q_y"], d["q_cb"], d["q_cr"]))


def test_golden_huffman_tables(oracle, golden):
    for name in _names(golden, "huff_names"):
        t = oracle.huffman_from_text(golden[f"huff/{name}/text"]).as_dict()
        for k in ("length", "code_msb", "counts", "symbols"):
            assert np.array_equal(t[k], golden[f"huff/{name}/{k}"]), (name, k)


def test_scan_is_decodable(oracle, golden):
    """independent check: PIL decodes the oracle's files to something close to the input"""
    PIL = pytest.importorskip("PIL.Image")
    import io
    for name in ("synth_128x128", "noise_96x80"):
        rgb, _ = oracle.ppm_load(golden[f"{name}/ppm"].tobytes())
        img = np.asarray(PIL.open(io.BytesIO(golden[f"{name}/jpg"].tobytes())).convert("RGB")).astype(np.float64)
        assert img.shape == rgb.shape
        err = np.abs(img - rgb).mean()
        assert err < (12 if name.startswith("synth") else 60), (name, err)


def test_golden_planes_files_byte_identical(oracle, golden_planes):
    """Image::writeJPEG on Images assembled from planes of doubles (edited samples, real-valued planes, YCbCr input): the
    oracle's restatement (jo_encode_planes) against the compiled reference's bytes"""
    for name in _names(golden_planes, "names"):
        planes = golden_planes[f"{name}/planes"]
        w, h, ycc = (int(x) for x in golden_planes[f"{name}/dims"])
        assert oracle.encode_planes(planes[0], planes[1], planes[2], w, h, bool(ycc)) == golden_planes[f"{name}/jpg"].tobytes(), name
