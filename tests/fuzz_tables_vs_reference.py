"""The product's host table builder (jpgenc_build_huffman) against the COMPILED REFERENCE's generateHuffmanCode
(oracle/_ref, i.e. only where /root/reference is mounted) on random symbol texts of five families:
    python tests/fuzz_tables_vs_reference.py <seed> <texts>
30 000 texts (seeds 1-6 x 5000): 0 mismatches."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ctypes as C
import numpy as np
from oracle.pyoracle import Oracle, Reference
from jpgenc_b200 import capi
lib = capi.load_library()
R = Reference(); o = Oracle()
seed = int(sys.argv[1]); trials = int(sys.argv[2])
rng = np.random.default_rng(seed)
bad = 0
for trial in range(trials):
    mode = trial % 5
    nsym = int(rng.integers(1, [6, 30, 120, 256, 60][mode] + 1))
    alphabet = rng.permutation(256)[:nsym]
    n = int(rng.integers(nsym, [200, 2000, 6000, 20000, 3000][mode]))
    if mode in (0, 4):
        text = alphabet[rng.integers(0, nsym, n)]
    else:
        p = rng.random(nsym) ** int(rng.integers(1, 9))
        text = rng.choice(alphabet, n, p=p / p.sum())
    r = R.huffman(text)
    count = np.bincount(text, minlength=256).astype(np.uint32)
    first = np.full(256, np.iinfo(np.uint64).max, np.uint64)
    idx = np.unique(text, return_index=True)
    first[idx[0]] = idx[1].astype(np.uint64) * 3
    tab = capi.HuffTable()
    assert lib.jpgenc_build_huffman(count.ctypes.data_as(capi.u32p), first.ctypes.data_as(capi.u64p), C.byref(tab)) == 0
    nn = int(sum(tab.counts))
    ok = (np.array_equal(np.array(tab.length), r["length"]) and np.array_equal(np.array(tab.code_msb), r["code_msb"])
          and np.array_equal(np.array(tab.counts), r["counts"]) and np.array_equal(np.array(tab.symbols[:nn]), r["symbols"][:nn]))
    if not ok:
        bad += 1
        if bad < 4: print("mismatch", trial, mode, nsym, n)
print(f"seed {seed}: {trials} texts, {bad} mismatches")
