"""CPU-side checks of the product library: it loads, exports every symbol the header declares, its host functions
(Huffman table build, header writer, PPM front end) agree with the oracle, and it refuses to run without a GPU.
No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, table_fields


def _lib():
    from jpgenc_b200 import capi
    return capi, capi.load_library()


def test_library_exports_every_declared_symbol():
    capi, lib = _lib()
    header = open(os.path.join(ROOT, "include", "jpgenc_b200.h")).read()
    declared = set(re.findall(r"\b(jpgenc_[a-z0-9_]+)\s*\(", header))
    declared -= {"jpgenc_ctx"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/jpgenc_b200.h but not exported"
    assert declared == set(capi._SIGNATURES), declared ^ set(capi._SIGNATURES)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    capi, lib = _lib()
    h = C.c_void_p()
    rc = lib.jpgenc_create(0, C.byref(h))
    assert rc == capi.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.jpgenc_last_error(None)
    with pytest.raises(capi.JpgencError):
        capi.Encoder(0)


def test_product_does_not_touch_the_oracle():
    """the shipped package must not import, link or execute anything under oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "jpgenc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "jo_" not in text, f
    text = open(os.path.join(ROOT, "Makefile")).read()
    assert "oracle" not in text


def _host_tables(lib, capi, count, first):
    tabs = (capi.HuffTable * 4)()
    for t in range(4):
        c = np.ascontiguousarray(count[t], np.uint32)
        f = np.ascontiguousarray(first[t], np.uint64)
        assert lib.jpgenc_build_huffman(c.ctypes.data_as(capi.u32p), f.ctypes.data_as(capi.u64p), C.byref(tabs[t])) == 0
    return tabs


def test_host_huffman_build_matches_oracle(oracle):
    capi, lib = _lib()
    rng = np.random.default_rng(21)
    for trial in range(200):
        nsym = int(rng.integers(1, 180))
        alphabet = rng.permutation(256)[:nsym]
        n = int(rng.integers(nsym, 3000))
        p = rng.random(nsym) ** (1 + trial % 4)
        text = rng.choice(alphabet, n, p=p / p.sum())
        count = np.bincount(text, minlength=256).astype(np.uint32)
        first = np.full(256, np.iinfo(np.uint64).max, np.uint64)
        for pos in range(n - 1, -1, -1):
            first[text[pos]] = pos * 64
        want = table_fields(oracle.huffman_from_text(text))
        for build in (lib.jpgenc_build_huffman, lib.jpgenc_build_huffman_containers):
            tab = capi.HuffTable()
            assert build(count.ctypes.data_as(capi.u32p), first.ctypes.data_as(capi.u64p), C.byref(tab)) == 0
            assert table_fields(tab) == want, trial


def test_host_huffman_golden(golden):
    capi, lib = _lib()
    for name in [str(n) for n in golden["huff_names"]]:
        text = golden[f"huff/{name}/text"]
        count = np.bincount(text, minlength=256).astype(np.uint32)
        first = np.full(256, np.iinfo(np.uint64).max, np.uint64)
        for pos in range(len(text) - 1, -1, -1):
            first[text[pos]] = pos
        tab = capi.HuffTable()
        assert lib.jpgenc_build_huffman(count.ctypes.data_as(capi.u32p), first.ctypes.data_as(capi.u64p), C.byref(tab)) == 0
        n = int(sum(tab.counts))
        assert np.array_equal(np.array(tab.length), golden[f"huff/{name}/length"])
        assert np.array_equal(np.array(tab.code_msb), golden[f"huff/{name}/code_msb"])
        assert np.array_equal(np.array(tab.symbols[:n]), golden[f"huff/{name}/symbols"])


def test_host_headers_match_oracle(oracle):
    capi, lib = _lib()
    from jpgenc_b200.synth import noise_rgb
    coef = oracle.forward(noise_rgb(48, 32, 2))
    count, first = oracle.symbol_stats(coef, 3, 2)
    tabs = _host_tables(lib, capi, count, first)
    otabs, _, _, _ = oracle.entropy_encode(coef, 3, 2)
    qy, qc = oracle.qy, oracle.qc
    n = lib.jpgenc_write_headers(48, 32, qy.ctypes.data_as(capi.u8p), qc.ctypes.data_as(capi.u8p), tabs, None)
    out = np.empty(n, np.uint8)
    lib.jpgenc_write_headers(48, 32, qy.ctypes.data_as(capi.u8p), qc.ctypes.data_as(capi.u8p), tabs, out.ctypes.data_as(capi.u8p))
    assert np.array_equal(out, oracle.headers(48, 32, otabs))


def _ppm_info(lib, data):
    w, h, m, off = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_size_t()
    magic = C.c_int()
    rc = lib.jpgenc_ppm_info(data, len(data), C.byref(w), C.byref(h), C.byref(m), C.byref(magic), C.byref(off))
    return rc, (magic.value, w.value, h.value, m.value, off.value)


@pytest.mark.parametrize("data", [
    b"P6\n4 2\n255\n" + bytes(range(24)),
    b"P6 4 2 255 " + bytes(range(24)),
    b"P6\n# a comment\n4 2\n# another\n255\n" + bytes(range(24)),
    b"P3\n2 2\n15\n1 2 3 4 5 6\n7 8 9 10 11 12\n",
    b"P3\n# c\n2 2 15 1 2 3 4 5 6 7 8 9 10 11 12",
    b"\n\n P6\t3\r\n1 63\n" + bytes(range(9)),
])
def test_ppm_front_end_matches_oracle(oracle, data):
    capi, lib = _lib()
    rc, info = _ppm_info(lib, data)
    orc, oh = oracle.ppm_parse(data)
    assert rc == 0 and orc == 0
    assert info == (oh.magic, oh.width, oh.height, oh.maxval, oh.payload)
    out = np.empty(info[1] * info[2] * 3, np.uint8)
    assert lib.jpgenc_ppm_samples(data, len(data), out.ctypes.data_as(capi.u8p)) == 0
    rgb, _ = oracle.ppm_load(data)
    assert np.array_equal(out, rgb.reshape(-1))


@pytest.mark.parametrize("data", [b"P5\n2 2\n255\n....", b"JFIF", b"P6\n2 2\n65535\n" + bytes(48), b"P6\n0 2\n255\n"])
def test_ppm_rejects_what_the_reference_rejects(data):
    capi, lib = _lib()
    rc, _ = _ppm_info(lib, data)
    assert rc == capi.ERR_FORMAT          # reference: runtime_error "Only P3 and P6 format is supported!" / assert on maxval


@pytest.mark.parametrize("data", [
    b"P6\n2 2\n63\n" + bytes([1, 2, 3, 4, 5, 6, 7, 8, 200, 9, 10, 11]),
    b"P3\n2 1\n100\n1 2 3 4 300 6\n",
    b"P3\n2 1\n15\n1 2 3 4 16 6\n",
])
def test_ppm_samples_above_maxval_are_refused(data):
    """the reference would scale such a sample past 255; the 8-bit device path is not exact for that, so the file is refused
    (jpgenc_encode_planes is the entry point for data of that kind)"""
    capi, lib = _lib()
    rc, info = _ppm_info(lib, data)
    assert rc == 0
    out = np.zeros(info[1] * info[2] * 3, np.uint8)
    assert lib.jpgenc_ppm_samples(data, len(data), out.ctypes.data_as(capi.u8p)) == capi.ERR_FORMAT


def test_golden_ppm_headers(golden, oracle):
    capi, lib = _lib()
    for name in [str(n) for n in golden["names"]]:
        data = golden[f"{name}/ppm"].tobytes()
        rc, info = _ppm_info(lib, data)
        orc, oh = oracle.ppm_parse(data)
        assert rc == 0 and info == (oh.magic, oh.width, oh.height, oh.maxval, oh.payload)


def test_array_restatement_of_the_table_build_equals_the_container_driven_build():
    """The device-side table build (csrc/tables_device.cu) restates libstdc++'s unordered_map iteration order and heap
    order on plain arrays.  Its code also runs on the host (jpgenc_build_huffman_arrays): every alphabet size, tie-heavy,
    geometric (length limit binds) and maximally deep weight patterns must give the tables of jpgenc_build_huffman_containers,
    which drives the real containers and is itself pinned against the reference (test_oracle_vs_reference) -- and so must
    jpgenc_build_huffman, the allocation-free build the encode path calls."""
    import ctypes as C
    import numpy as np
    from conftest import random_symbol_stats, table_fields
    from jpgenc_b200.capi import HuffTable, _np_ptr, u32p, u64p
    _, lib = _lib()
    rng = np.random.default_rng(7)
    sizes = list(range(1, 70)) + [100, 127, 128, 129, 161, 162, 200, 255, 256] + [int(x) for x in rng.integers(1, 257, 700)]
    for n in sizes:
        c, f = random_symbol_stats(rng, n, int(rng.choice([5, 1000, 10 ** 6, 2 ** 31 - 1])))
        a, b, p = HuffTable(), HuffTable(), HuffTable()
        assert lib.jpgenc_build_huffman_containers(_np_ptr(c, u32p), _np_ptr(f, u64p), C.byref(a)) == 0
        assert lib.jpgenc_build_huffman_arrays(_np_ptr(c, u32p), _np_ptr(f, u64p), C.byref(b)) == 0
        assert lib.jpgenc_build_huffman(_np_ptr(c, u32p), _np_ptr(f, u64p), C.byref(p)) == 0
        assert table_fields(a) == table_fields(b) and a.nsymbols == b.nsymbols, f"{n} symbols: {c[c > 0].tolist()}"
        assert table_fields(a) == table_fields(p) and a.nsymbols == p.nsymbols, f"{n} symbols: {c[c > 0].tolist()}"
    # a count past INT_MAX wraps in the reference's int counter: the packed build hands such input to the plain one
    c, f = random_symbol_stats(rng, 9, 1000)
    c[np.flatnonzero(c)[0]] = 2 ** 31 + 5
    a, p = HuffTable(), HuffTable()
    assert lib.jpgenc_build_huffman_containers(_np_ptr(c, u32p), _np_ptr(f, u64p), C.byref(a)) == lib.jpgenc_build_huffman(_np_ptr(c, u32p), _np_ptr(f, u64p), C.byref(p))
    assert table_fields(a) == table_fields(p)
    empty = np.zeros(256, np.uint32)
    none = np.full(256, np.iinfo(np.uint64).max, np.uint64)
    assert lib.jpgenc_build_huffman_arrays(_np_ptr(empty, u32p), _np_ptr(none, u64p), C.byref(HuffTable())) != 0


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """bench.py --impl reference (the CPU arm the driver runs beside ours) needs no GPU: exactly one JSON line on stdout
    with the contract's keys, whatever the libraries print while loading"""
    import json
    import subprocess
    import sys
    env = dict(os.environ, JPGENC_BENCH_REF_SIZE="512x512")        # the real arm encodes the 16384x16384 file itself (~1 min)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encode_throughput" and d["unit"] == "Mpx/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and d["cpu_baseline"]["cores"] == 1
    # same `config` object as our arm prints for this workload
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.headline_config("image16k", 1)


@pytest.mark.parametrize("workers", [1, 2, 3])
def test_table_pool_handover_under_stress(tmp_path_factory, workers):
    """TablePool (csrc/host_pools.hpp) hands the four histograms of an image to its worker thread(s) through an armed /
    published / done handshake with spinning waits: thousands of cycles on random, tie-heavy histograms, the release-without-
    work path in between, every table compared with a direct build (tests/host/table_pool_probe.cu; clean under
    -fsanitize=thread as well)"""
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    libdir = os.path.join(ROOT, "jpgenc_b200", "lib")
    exe = str(tmp_path_factory.getbasetemp() / "table_pool_probe")
    if not os.path.exists(exe):
        subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                        os.path.join(ROOT, "tests", "host", "table_pool_probe.cu"), "-o", exe, "-L" + libdir, "-ljpgenc_b200",
                        "-Xlinker", "-rpath", "-Xlinker", libdir], check=True)
    r = subprocess.run([exe, str(workers), "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_host_pool_parallel_for_under_stress(tmp_path_factory):
    """HostPool::parallel_for with several callers at once on one pool (the lanes of a threaded batch, the band reader of
    jpgenc_encode_ppm_file): every job runs exactly once, a call returns only after its last job
    (tests/host/host_pool_probe.cu; clean under -fsanitize=thread as well)"""
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    libdir = os.path.join(ROOT, "jpgenc_b200", "lib")
    exe = str(tmp_path_factory.getbasetemp() / "host_pool_probe")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                    os.path.join(ROOT, "tests", "host", "host_pool_probe.cu"), "-o", exe, "-L" + libdir, "-ljpgenc_b200",
                    "-Xlinker", "-rpath", "-Xlinker", libdir], check=True)
    for workers, callers in ((3, 4), (1, 6), (7, 2)):
        r = subprocess.run([exe, str(workers), str(callers), "1000"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
