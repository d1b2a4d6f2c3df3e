"""The C++ host mirror (jpgenc_b200/host) seen from a caller of the reference's API.

CPU: the stage methods of `Image` (convertToColorSpace .. doHuffmanEncoding), the Huffman/Bitstream/Segment classes
reproduce the compiled reference's golden vectors bit for bit; where /root/reference is mounted, the REFERENCE'S OWN
unit tests (src/test/*.cpp, unchanged) are compiled against the mirror headers and must behave exactly as they do
against the reference's own headers on this compiler (the 10 MSVC-ordering checks of HuffmanTest/PackageMergeTest fail
identically on both).
GPU: Image::writeJPEG / the jpgEnc command line produce byte-identical files.
"""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
INC = os.path.join(ROOT, "jpgenc_b200", "host", "include")
LIBDIR = os.path.join(ROOT, "jpgenc_b200", "lib")
REF = "/root/reference"
LINK = ["-L" + LIBDIR, "-ljpgenc_b200", "-Wl,-rpath," + LIBDIR]


def _need_lib():
    if not os.path.exists(os.path.join(LIBDIR, "libjpgenc_b200.so")):
        pytest.fail("libjpgenc_b200.so not built: run `make` (or __graft_entry__.build())")


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    _need_lib()
    exe = str(tmp_path_factory.mktemp("mirror") / "mirror_probe")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-access-control", "-I" + INC,
                    os.path.join(ROOT, "tests", "host", "mirror_probe.cpp"), "-o", exe] + LINK, check=True)
    return exe


def _read_dump(path):
    raw = open(path, "rb").read()
    out, at = [], 0
    for dtype in (np.float64, np.float64, np.float64, np.int32, np.int32, np.int32, np.uint8):
        r, c = struct.unpack_from("<II", raw, at)
        at += 8
        n = r * c * np.dtype(dtype).itemsize
        out.append(np.frombuffer(raw, dtype, r * c, at).reshape(r, c))
        at += n
    assert at == len(raw)
    return out


def test_stage_methods_match_reference_golden(probe, golden, tmp_path):
    for name in [str(n) for n in golden["names"]]:
        ppm = tmp_path / (name + ".ppm")
        ppm.write_bytes(golden[f"{name}/ppm"].tobytes())
        dump = tmp_path / (name + ".bin")
        subprocess.run([probe, "stages", str(ppm), str(dump)], check=True, capture_output=True)
        y, cb, dct_y, q_y, q_cb, q_cr, scan = _read_dump(dump)
        assert np.array_equal(q_y, golden[f"{name}/q_y"]), name
        assert np.array_equal(q_cb, golden[f"{name}/q_cb"]), name
        assert np.array_equal(q_cr, golden[f"{name}/q_cr"]), name
        if f"{name}/y" in golden:
            assert np.array_equal(y, golden[f"{name}/y"]), name                 # doubles, bit for bit
            assert np.array_equal(cb, golden[f"{name}/cb"]), name
            assert np.array_equal(dct_y, golden[f"{name}/dct_y"]), name
        # the stage API's own scan == the entropy-coded segment inside the reference's file
        jpg = golden[f"{name}/jpg"].tobytes()
        sos = jpg.index(b"\xff\xda")
        start = sos + 2 + struct.unpack(">H", jpg[sos + 2:sos + 4])[0]
        assert scan.tobytes() == jpg[start:-2], name


def test_huffman_adaptor_matches_table_builder(probe):
    r = subprocess.run([probe, "huffman"], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout + r.stderr


def test_segment_classes_match_header_writer(probe):
    r = subprocess.run([probe, "segments"], capture_output=True, text=True)
    assert r.returncode == 0 and "identical 1" in r.stdout, r.stdout + r.stderr


def test_write_jpeg_without_gpu_throws(probe, golden, tmp_path):
    """no CPU fallback: on a machine without a B200 the hot path fails loudly (std::runtime_error -> exit 3)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ppm = tmp_path / "a.ppm"
    ppm.write_bytes(golden["synth_64x48/ppm"].tobytes())
    r = subprocess.run([probe, "encode", str(ppm), str(tmp_path / "a.jpg")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr, r.stderr
    assert not (tmp_path / "a.jpg").exists() or (tmp_path / "a.jpg").stat().st_size == 0


# ---- the reference's own unit tests, unchanged, against the mirror ------------------------------------------------
REF_TESTS = ["DctTest", "CodingTest", "BitstreamGenericTest", "HuffmanTest", "PackageMergeTest", "ImageTest"]
SHIM = os.path.join(ROOT, "tests", "host", "boost_shim")          # Boost.Test stand-in
UBLAS_FWD = os.path.join(ROOT, "tests", "host", "ublas_fwd")      # <boost/numeric/ublas/*.hpp> -> the mirror's matrix types


def _build_ref_test(name, tmp, against_mirror):
    src = [os.path.join(REF, "src", "test", name + ".cpp"), os.path.join(REF, "src", "test", "common.cpp")]
    exe = os.path.join(tmp, name + ("_mirror" if against_mirror else "_ref"))
    if against_mirror:
        cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-include", "algorithm", "-include", "cmath",
               "-I" + SHIM, "-I" + UBLAS_FWD, "-I" + INC, "-I" + os.path.join(REF, "include")] + src + ["-o", exe] + LINK
    else:
        # the reference's headers need the three g++ fixes of oracle/build_ref.sh: patch a throw-away copy
        inc = os.path.join(tmp, "ref_include")
        if not os.path.isdir(inc):
            shutil.copytree(os.path.join(REF, "include"), inc)
            subprocess.run(["sed", "-i", "-E", r"/^    template<typename BlockType>$/{N;s/BlockType/BlockTypeF/g}",
                            os.path.join(inc, "BitstreamGeneric.hpp")], check=True)
            subprocess.run(["sed", "-i", "-E", r"s/template <int T>/template <std::size_t T>/; s/template <int sz>/template <std::size_t sz>/; "
                            r"s/HTinfo\.assign\(/HTinfo.fill(/; s/QT\.QT_info\.assign\(/QT.QT_info.fill(/",
                            os.path.join(inc, "JpegSegments.hpp")], check=True)
        cmd = ["g++", "-std=c++20", "-O1", "-fopenmp", "-w", "-include", "algorithm", "-include", "cmath", "-include", "functional",
               "-include", "array", "-I" + SHIM, "-I" + os.path.join(ROOT, "oracle", "boost_shim"), "-I" + inc] + src + \
              [os.path.join(REF, "src", "Image.cpp"), os.path.join(REF, "src", "Huffman.cpp"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def _verdicts(exe, cwd, skip):
    env = dict(os.environ, JPGENC_SKIP_CASES=skip, OMP_NUM_THREADS="1")
    r = subprocess.run([exe], cwd=cwd, capture_output=True, env=env)
    lines = [l for l in (r.stdout + r.stderr).decode("latin-1").splitlines()
             if l.startswith(("ok ", "FAILED", "skipped")) or "error in" in l or " cases, " in l]
    return sorted(l.split("/")[-1] for l in lines)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "test")), reason="/root/reference not mounted")
@pytest.mark.parametrize("name", REF_TESTS)
def test_reference_unit_tests_behave_identically_on_the_mirror(name, tmp_path):
    _need_lib()
    tmp = str(tmp_path)
    os.symlink(os.path.join(REF, "src", "test", "res"), os.path.join(tmp, "res"))
    skip = "jpeg_segments_test"          # calls writeJPEG: needs the GPU (covered by the gpu tests below)
    mirror = _verdicts(_build_ref_test(name, tmp, True), tmp, skip)
    ref = _verdicts(_build_ref_test(name, tmp, False), tmp, skip)
    assert mirror == ref
    assert any(l.startswith("ok ") for l in mirror)
    if name not in ("HuffmanTest", "PackageMergeTest"):
        assert not any(l.startswith("FAILED") for l in mirror), mirror


# ---- GPU: the hot path through the C++ surface ----------------------------------------------------------------------
@pytest.mark.gpu
def test_write_jpeg_byte_identical(probe, golden, tmp_path):
    for name in [str(n) for n in golden["names"]]:
        ppm, jpg = tmp_path / (name + ".ppm"), tmp_path / (name + ".jpg")
        ppm.write_bytes(golden[f"{name}/ppm"].tobytes())
        subprocess.run([probe, "encode", str(ppm), str(jpg)], check=True, capture_output=True)
        assert jpg.read_bytes() == golden[f"{name}/jpg"].tobytes(), name


@pytest.mark.gpu
def test_stage_methods_on_the_device_match_reference_golden(probe, golden, tmp_path):
    """Image::stagesOnDevice(true): applySubsampling / applyDCT run on the GPU (jpgenc_stage_subsample / jpgenc_stage_dct) inside the
    reference's stage sequence; every plane that follows must still be the compiled reference's"""
    for name in [str(n) for n in golden["names"]]:
        ppm = tmp_path / (name + ".ppm")
        ppm.write_bytes(golden[f"{name}/ppm"].tobytes())
        dump = tmp_path / (name + ".bin")
        subprocess.run([probe, "stages_device", str(ppm), str(dump)], check=True, capture_output=True)
        y, cb, dct_y, q_y, q_cb, q_cr, scan = _read_dump(dump)
        assert np.array_equal(q_y, golden[f"{name}/q_y"]), name
        assert np.array_equal(q_cb, golden[f"{name}/q_cb"]) and np.array_equal(q_cr, golden[f"{name}/q_cr"]), name
        if f"{name}/y" in golden:
            assert np.array_equal(cb, golden[f"{name}/cb"]) and np.array_equal(dct_y, golden[f"{name}/dct_y"]), name   # doubles, exact


@pytest.mark.gpu
def test_write_jpeg_encodes_the_planes_not_the_file(probe, golden, oracle, tmp_path):
    """Image::writeJPEG encodes what R/G/B hold when it is called (src/Image.cpp:831-846) and converts only an RGB image
    (:112-115, 839): planes edited after loadPPM, a real-valued sample, an edited padding sample, an image converted to
    YCbCr by the caller"""
    for name in ("grad_26x19", "synth_64x48", "p6_max63"):
        ppm, jpg = tmp_path / (name + ".ppm"), tmp_path / (name + ".jpg")
        ppm.write_bytes(golden[f"{name}/ppm"].tobytes())
        rgb, maxval = oracle.ppm_load(golden[f"{name}/ppm"].tobytes())
        h, w, _ = rgb.shape
        pad = np.pad(rgb, ((0, -h % 16), (0, -w % 16), (0, 0)), mode="edge").astype(np.float64) * (255. / maxval)
        # already YCbCr: the same file as the plain encode (the conversion happened earlier, with the same arithmetic)
        subprocess.run([probe, "encode_ycc", str(ppm), str(jpg)], check=True, capture_output=True)
        assert jpg.read_bytes() == golden[f"{name}/jpg"].tobytes(), name
        # edited, still 8-bit samples (takes the 8-bit path again, with the new samples)
        p = pad.copy()
        p[1, 2, 0] = 255; p[0, 0, 1] = 0; p[2, 1, 2] = 17
        subprocess.run([probe, "encode_edited8", str(ppm), str(jpg)], check=True, capture_output=True)
        assert jpg.read_bytes() == oracle.encode_planes(p[..., 0], p[..., 1], p[..., 2], w, h), name
        # edited to something that is not an 8-bit image: the planes go to the device as doubles
        p = pad.copy()
        p[1, 2, 0] += 0.37
        p[-1, -1, 2] = 99.5
        subprocess.run([probe, "encode_edited", str(ppm), str(jpg)], check=True, capture_output=True)
        assert jpg.read_bytes() == oracle.encode_planes(p[..., 0], p[..., 1], p[..., 2], w, h), name


@pytest.mark.gpu
def test_command_line_like_the_reference(golden, tmp_path):
    """jpgEnc <in.ppm> [out.jpg] (reference src/main.cpp:8-32): default output name, error on a missing file"""
    exe = os.path.join(ROOT, "jpgenc_b200", "bin", "jpgEnc")
    assert os.path.exists(exe), "run `make`"
    ppm = tmp_path / "in.ppm"
    ppm.write_bytes(golden["noise_96x80/ppm"].tobytes())
    r = subprocess.run([exe, str(ppm)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "noname.jpg").read_bytes() == golden["noise_96x80/jpg"].tobytes()
    assert "Encoding duration" in r.stdout and "PPM loading took" in r.stdout
    r = subprocess.run([exe, str(tmp_path / "missing.ppm"), str(tmp_path / "x.jpg")], capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to open" in r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert "No filename" in r.stdout
