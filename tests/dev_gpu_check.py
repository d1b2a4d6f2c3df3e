"""Stage-by-stage GPU-vs-oracle diagnostics (development aid, run from the repository root: python tests/dev_gpu_check.py; the
parity tests proper are the test_*.py files).  Lives under tests/ because it uses the oracle."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from jpgenc_b200.capi import Encoder
from jpgenc_b200.synth import synth_rgb, noise_rgb
from oracle.pyoracle import Oracle

O = Oracle()
E = Encoder(0)
E.set_stage_timing(2)


def check_image(name, rgb, maxval=255):
    h, w, _ = rgb.shape
    w16, h16, mw, mh = O.geometry(w, h)
    print(f"== {name} {w}x{h} maxval={maxval} mcu={mw}x{mh}")
    ref = O.forward(rgb, maxval)
    E.upload_rgb(rgb, maxval)
    E.color_dct_quant()
    got = E.get_coefficients()
    st = E.stats()
    diff = (got.astype(np.int32) - ref.astype(np.int32))
    nbad = int(np.count_nonzero(diff))
    print(f"  K1: coeff mismatches {nbad}/{ref.size} max|d|={int(np.abs(diff).max())} refined={st.refined_blocks}/{st.n_blocks} ms_forward={st.ms_forward:.3f}")
    if nbad:
        idx = np.argwhere(diff != 0)[:10]
        for m, k, z in idx:
            print(f"     mcu {m} blk {k} zz {z}: got {got[m,k,z]} ref {ref[m,k,z]}")
    # K2 on the reference coefficients so that later stages are independent of K1
    E.set_coefficients_mcu(ref, mw, mh)
    cnt, first = E.symbol_stats()
    ocnt, ofirst = O.symbol_stats(ref, mw, mh)
    print(f"  K2: hist equal {np.array_equal(cnt, ocnt)} first_pos equal {np.array_equal(first, ofirst)} ms_stats={E.stats().ms_stats:.3f}")
    if not np.array_equal(cnt, ocnt):
        bad = np.argwhere(cnt != ocnt)[:10]
        for t, s in bad: print(f"     table {t} sym {s:02x}: got {cnt[t,s]} ref {ocnt[t,s]}")
    if not np.array_equal(first, ofirst):
        bad = np.argwhere(first != ofirst)[:10]
        for t, s in bad: print(f"     first table {t} sym {s:02x}: got {first[t,s]} ref {ofirst[t,s]}")
    tabs = E.build_huffman(ocnt, ofirst)
    otabs, oraw, onbits, ostuffed = O.entropy_encode(ref, mw, mh)
    ok_t = all(bytes(tabs[t].length) == bytes(otabs[t].length) and bytes(tabs[t].code_msb) == bytes(otabs[t].code_msb)
               and bytes(tabs[t].counts) == bytes(otabs[t].counts) and bytes(tabs[t].symbols) == bytes(otabs[t].symbols) for t in range(4))
    print(f"  host tables equal {ok_t}")
    try:
        n = E.entropy_encode(tabs)
        scan = E.download_scan()
        st = E.stats()
        print(f"  K3/K4: scan bytes {n} ref {ostuffed.size} bits {st.scan_bits} ref {onbits} ff {st.stuffed_ff} equal {np.array_equal(scan, ostuffed)} ms_entropy={st.ms_entropy:.3f}")
        if not np.array_equal(scan, ostuffed):
            m = min(scan.size, ostuffed.size)
            d = np.nonzero(scan[:m] != ostuffed[:m])[0]
            print(f"     first diffs at {d[:10]} of {d.size}; got {scan[d[:8]]} ref {ostuffed[d[:8]]}")
    except Exception as ex:
        print("  K3/K4 FAILED:", ex)
    # whole file
    try:
        mine = E.encode_rgb(rgb, maxval)
        import tempfile, os
        if maxval == 255:
            theirs = O.encode_rgb(rgb)
            print(f"  file: {len(mine)} bytes ref {len(theirs)} identical {mine == theirs}")
    except Exception as ex:
        print("  whole-file FAILED:", ex)


def main():
    try:
        check_image("synth512", synth_rgb(512, 512, 0))
        check_image("noise256", noise_rgb(256, 256, 1))
        check_image("odd", synth_rgb(203, 117, 3))
        check_image("tiny", synth_rgb(5, 3, 1))
        check_image("noise-odd", noise_rgb(100, 37, 2))
        check_image("maxval15", (noise_rgb(64, 48, 5) >> 4), 15)
        check_image("flat", np.full((64, 64, 3), 200, np.uint8))
        check_image("synth1080", synth_rgb(1920, 1080, 0))
    except Exception:
        traceback.print_exc()
    # microbench
    try:
        nb = 1 << 16
        d_in = E.dev_alloc(nb * 256); d_out = E.dev_alloc(nb * 128)
        E.synth_blocks(d_in, nb)
        refined = E.dct_quant_blocks(d_in, d_out, nb, O.qy)
        x = np.empty((nb, 64), np.float32); E.d2h(x, d_in)
        y = np.empty((nb, 64), np.int16); E.d2h(y, d_out)
        zz = np.array([O.zigzag_index(i) for i in range(64)])
        bad = 0
        for b in range(0, nb, 97):
            q = O.quantize(O.dct(x[b].astype(np.float64).reshape(8, 8)), O.qy).reshape(64)[zz]
            bad += int(np.count_nonzero(q != y[b]))
        print(f"== microbench {nb} blocks: refined {refined} mismatching coeffs on sample {bad}")
        E.dev_free(d_in); E.dev_free(d_out)
    except Exception:
        traceback.print_exc()
    # timing at scale
    for (w, h) in [(3840, 2160), (16384, 16384)]:
        try:
            d = E.dev_alloc(w * h * 3)
            E.synth_rgb(d, w, h, 0)
            E.bind_device_rgb(d, w, h)
            for it in range(3):
                t = time.time(); n = E.encode_bound(None); E.synchronize(); dt = time.time() - t
                st = E.stats()
                print(f"== {w}x{h}: jpeg {n} B wall {dt*1e3:.2f} ms  K1 {st.ms_forward:.3f} K2 {st.ms_stats:.3f} K3+4 {st.ms_entropy:.3f} refined {st.refined_blocks} -> {w*h/dt/1e6:.0f} Mpx/s")
            E.dev_free(d)
        except Exception:
            traceback.print_exc()


main()
