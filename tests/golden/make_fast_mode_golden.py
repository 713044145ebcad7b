#!/usr/bin/env python
"""Regenerate tests/golden/fast_mode_outputs.json: what the engine's FAST-mode parses (the scalar restatement in
oracle/cw_oracle.c of compute_war_b200/csrc/lz4_fast.cuh) produce over tests/golden/corpus_sample.bin -- compressed
lengths and a SHA-256 of every block, per file and chunk size, for LZ4 and LZF.

    python tests/golden/make_fast_mode_golden.py        (needs only oracle/liboracle.so and the sample; no reference tree)

The fast parse is not a reference algorithm (the reference constrains it through its decoder and its size only); this
fixture pins it across versions: `-m "not gpu"` checks that the restatement still produces these blocks, `-m gpu` that the
kernels do.  A deliberate change of the parse rules regenerates it (and re-runs the ratio / reference-decoder tests).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def main():
    ref = json.load(open(os.path.join(HERE, "ref_outputs.json")))
    blob = np.fromfile(os.path.join(HERE, "corpus_sample.bin"), np.uint8)
    out = {"_source": "oracle/cw_oracle.c cwo_lz4_fast_compress / cwo_lzf_fast_compress over corpus_sample.bin", "corpus": {}}
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        per = {}
        for chunk in (4096, 8192, 65536):
            n = len(d) // chunk
            if n == 0:
                continue
            rec = {"n": n}
            for name, fn in (("lz4", O.lz4_fast_chunks), ("lzf", O.lzf_fast_chunks)):
                blocks, lens = fn(d, chunk)
                rec[name + "_lens"] = [int(x) for x in lens]
                rec[name + "_sha256"] = [hashlib.sha256(blocks[i, : lens[i]].tobytes()).hexdigest()[:32] for i in range(n)]
            per[str(chunk)] = rec
        out["corpus"][ent["name"]] = per
    json.dump(out, open(os.path.join(HERE, "fast_mode_outputs.json"), "w"), indent=0)
    print("wrote fast_mode_outputs.json")


if __name__ == "__main__":
    main()
