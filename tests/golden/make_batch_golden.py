"""Golden sizes/hashes of BASELINE config 4 (1024 synthetic 1920x1080 frames, frame k = seed k), produced by the CPU oracle
(which is byte-identical to the compiled reference on this generator: tests/test_oracle_vs_reference.py).

    python tests/golden/make_batch_golden.py        # writes tests/golden/batch1080p.json (takes a few minutes)

bench.py and the GPU tests compare the CUDA path's per-frame sizes and the SHA-256 of the concatenated files against it
without needing the oracle at run time."""
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

N, W, H = 1024, 1920, 1080


def one(k):
    from jpgenc_b200.synth import synth_rgb
    from oracle.pyoracle import Oracle
    jpg = Oracle().encode_rgb(synth_rgb(W, H, k))
    return k, len(jpg), hashlib.sha256(jpg).hexdigest()


def main():
    with ProcessPoolExecutor(os.cpu_count() or 4) as ex:
        rows = sorted(ex.map(one, range(N), chunksize=8))
    sizes = [r[1] for r in rows]
    all_hash = hashlib.sha256("".join(r[2] for r in rows).encode()).hexdigest()
    out = {"frames": N, "width": W, "height": H, "generator": "jpgenc_b200.synth.synth_rgb(W, H, seed=k)",
           "sizes": sizes, "total_bytes": sum(sizes), "sha256_frame0": rows[0][2],
           "sha256_of_frame_sha256s": all_hash,
           "note": "sha256_of_frame_sha256s = sha256 of the concatenated lowercase hex digests of the 1024 files, in frame order"}
    with open(os.path.join(ROOT, "tests", "golden", "batch1080p.json"), "w") as f:
        json.dump(out, f)
    print(out["total_bytes"], out["sha256_frame0"], all_hash)


if __name__ == "__main__":
    main()
