#!/usr/bin/env python
"""Generate tests/golden/golden.json from the REFERENCE's own codec.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
It compiles /root/reference/cbits/lz4.c unmodified (oracle/Makefile -> oracle/_ref/libreflz4.so),
drives it with the Haskell shim's call sequence (oracle/ref_driver.c) on seeded inputs
(streamly_lz4_b200/datagen, deterministic) and records, per case, the generator parameters,
every block's compressed length, the SHA-256 of the framed stream and -- for the small cases --
the framed bytes themselves (hex).  The reference ships no golden vectors of its own
(SURVEY.md section 8c), so these outputs of the reference run here are the pin.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle, available  # noqa: E402
from streamly_lz4_b200 import datagen  # noqa: E402

CASES = [
    # name, kind, seed, total bytes, block size, accel, linked, block_size config
    ("text_64k_linked_a1", "text", 101, 1 << 20, 65536, 1, True, "BlockHasSize"),
    ("text_64k_indep_a1", "text", 101, 1 << 20, 65536, 1, False, "BlockHasSize"),
    ("mixed_640000_indep_a400", "mixed", 2, 4 * 640000, 640000, 400, False, "BlockHasSize"),
    ("mixed_640000_indep_a1", "mixed", 2, 4 * 640000, 640000, 1, False, "BlockHasSize"),
    ("mixed_64k_linked_a1", "mixed", 7, 1 << 20, 65536, 1, True, "BlockHasSize"),
    ("records_4k_linked_a5", "records", 9, 1 << 18, 4096, 5, True, "BlockHasSize"),
    ("sparse_100000_linked_a12", "sparse01", 4, 1 << 20, 100000, 12, True, "BlockHasSize"),
    ("bits01_10k_linked_a-1", "bits01", 5, 200000, 10000, -1, True, "BlockHasSize"),
    ("biased01_100k_linked_a100", "biased01", 6, 500000, 100000, 100, True, "BlockHasSize"),
    ("random_64k_indep_a1", "random", 8, 1 << 18, 65536, 1, False, "BlockHasSize"),
    ("mixed_256k_max256_a1", "mixed", 12, 1 << 20, 200000, 1, True, "BlockMax256KB"),
    ("text_65547_edge", "text", 13, 65547 * 2, 65547, 1, False, "BlockHasSize"),
    ("mixed_a65537", "mixed", 14, 2 * 640000, 640000, 65537, False, "BlockHasSize"),
    ("tiny_text_300", "text", 15, 300, 300, 1, False, "BlockHasSize"),
    ("tiny_bits_1000_b100", "bits01", 16, 1000, 100, 1, True, "BlockHasSize"),
    ("tiny_sizes", "text", 17, 0, 0, 1, True, "BlockHasSize"),        # special: explicit size list
]
TINY_SIZES = [0, 1, 4, 12, 13, 14, 0, 3, 64, 2, 500, 0, 15, 31]


def arrays_of(case):
    name, kind, seed, total, bs, accel, linked, cfg = case
    if name == "tiny_sizes":
        d = datagen.make(kind, seed, 4096)
        out, at = [], 0
        for n in TINY_SIZES:
            out.append(d[at:at + n].tobytes()); at += n
        return out
    d = datagen.make(kind, seed, total)
    return [d[i:i + bs].tobytes() for i in range(0, total, bs)]


def main():
    assert available("reference") or os.path.exists("/root/reference/cbits/lz4.c"), "needs the reference tree"
    ref = Oracle("reference")
    out = {"generator": "tests/golden/make_golden.py", "codec": "reference cbits/lz4.c (LZ4 1.9.3), oracle/ref_driver.c call sequence",
           "cases": []}
    for case in CASES:
        name, kind, seed, total, bs, accel, linked, cfg = case
        arrays = arrays_of(case)
        framed = ref.compress_chunks(arrays, accel, block_size=cfg, linked=linked)
        blob = b"".join(framed)
        entry = {"name": name, "kind": kind, "seed": seed, "total": total, "block": bs, "accel": accel,
                 "linked": linked, "block_size": cfg, "input_sha256": hashlib.sha256(b"".join(arrays)).hexdigest(),
                 "framed_lens": [len(f) for f in framed], "framed_sha256": hashlib.sha256(blob).hexdigest()}
        if len(blob) <= 4096:
            entry["framed_hex"] = blob.hex()
            entry["input_hex"] = b"".join(arrays).hex()
        out["cases"].append(entry)
        print(name, len(arrays), "blocks", len(blob), "bytes")
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
