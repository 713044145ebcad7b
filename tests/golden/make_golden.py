#!/usr/bin/env python
"""Regenerate tests/golden/* from the reference tree.  Run in the BUILD container only:

    make -C oracle ref && python tests/golden/make_golden.py

It reads (never copies source from) /root/reference:
  * reference_code/skein/KAT_MCT/*.txt         -> skein_kat.json   (the reference's own golden vectors)
  * dataset/canterbury*, via oracle/_ref/corpus -> corpus_sample.bin (first 64 KiB of every corpus file;
                                                   data files, not source)
  * oracle/_ref/ref_cli (the reference's Skein source + its vendored liblz4.a 1.8.2 / liblzf.a +
    OpenSSL SHA256) run on those bytes          -> ref_outputs.json (digests, compressed lengths and a
                                                   SHA-256 of every compressed block)
The fixtures are what `-m "not gpu"` pins the oracle against and what `-m gpu` pins the CUDA path against on
the GPU box, where /root/reference does not exist.
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

KAT = "/root/reference/reference_code/skein/KAT_MCT"
SAMPLE_BYTES = 65536
CHUNKS = (4096, 8192, 65536)


def nist_file(name, keep):
    out = []
    txt = open(os.path.join(KAT, name)).read()
    for m in re.finditer(r"Len = (\d+)\s+Msg = ([0-9A-Fa-f]+)\s+MD = ([0-9A-Fa-f]+)", txt):
        bits = int(m.group(1))
        if keep(bits):
            msg = m.group(2).lower()
            if bits == 0:
                msg = ""
            out.append([bits, msg, m.group(3).lower()])
    return out


def golden_file():
    """sequential, un-keyed cases of skein_golden_kat.txt for Skein-256 and Skein-512"""
    out = []
    txt = open(os.path.join(KAT, "skein_golden_kat.txt")).read()
    for case in txt.split("--------------------------------"):
        h = re.search(r":Skein-(\d+):\s+(\d+)-bit hash, msgLen =\s*(\d+) bits, data = '([^']+)'", case)
        if not h:
            continue  # tree cases have a different header
        state, hbits, mbits, kind = int(h.group(1)), int(h.group(2)), int(h.group(3)), h.group(4)
        if state == 1024 or "MAC" in kind or "Tree" in case:
            continue
        body = case[h.end():]
        msg_part, res_part = body.split("Result:")
        msg_hex = "" if "(none)" in msg_part else "".join(re.findall(r"\b[0-9A-F]{2}\b", msg_part.split("Message data:")[1]))
        res_hex = "".join(re.findall(r"\b[0-9A-F]{2}\b", res_part))
        out.append([state, hbits, mbits, msg_hex.lower(), res_hex.lower()])
    return out


def monte_carlo(name):
    txt = open(os.path.join(KAT, name)).read()
    seed = re.search(r"Seed = ([0-9A-F]+)", txt).group(1).lower()
    mds = [m.lower() for m in re.findall(r"MD = ([0-9A-F]+)", txt)]
    return {"seed": seed, "md": mds}


def main():
    short = lambda b: b % 8 == 0 or b <= 1032  # noqa: E731
    aligned = lambda b: b % 8 == 0  # noqa: E731
    kat = {
        "_source": "reference_code/skein/KAT_MCT (NIST round-1 Skein v1.1 vectors); state size follows "
                   "SHA3api_ref.c:21-39 (<=256-bit digests -> Skein-256, <=512 -> Skein-512)",
        "ShortMsgKAT_512": {"state": 512, "hash": 512, "cases": nist_file("ShortMsgKAT_512.txt", short)},
        "ShortMsgKAT_384": {"state": 512, "hash": 384, "cases": nist_file("ShortMsgKAT_384.txt", aligned)},
        "ShortMsgKAT_256": {"state": 256, "hash": 256, "cases": nist_file("ShortMsgKAT_256.txt", short)},
        "ShortMsgKAT_224": {"state": 256, "hash": 224, "cases": nist_file("ShortMsgKAT_224.txt", aligned)},
        "LongMsgKAT_512": {"state": 512, "hash": 512, "cases": nist_file("LongMsgKAT_512.txt", aligned)},
        "LongMsgKAT_256": {"state": 256, "hash": 256, "cases": nist_file("LongMsgKAT_256.txt", aligned)},
        "MonteCarlo_512": dict(state=512, hash=512, **monte_carlo("MonteCarlo_512.txt")),
        "MonteCarlo_256": dict(state=256, hash=256, **monte_carlo("MonteCarlo_256.txt")),
        "ExtremelyLong_512": {
            "state": 512, "hash": 512, "repeat": 16777216,
            "text": "abcdefghbcdefghicdefghijdefghijkefghijklfghijklmghijklmnhijklmno",
            "md": re.search(r"MD = ([0-9A-F]+)", open(os.path.join(KAT, "ExtremelyLongMsgKAT_512.txt")).read()).group(1).lower(),
        },
        "golden": golden_file(),
    }
    with open(os.path.join(HERE, "skein_kat.json"), "w") as f:
        json.dump(kat, f, separators=(",", ":"))
    print("skein_kat.json:", {k: (len(v["cases"]) if isinstance(v, dict) and "cases" in v else "-") for k, v in kat.items()
                              if k != "_source"}, "golden", len(kat["golden"]))

    # corpus sample
    files = O.corpus_files()
    assert files, "run `make -C oracle ref` first"
    index, blob = [], bytearray()
    for p in files:
        d = open(p, "rb").read()[:SAMPLE_BYTES]
        index.append({"name": os.path.relpath(p, O.CORPUS_DIR), "offset": len(blob), "size": len(d),
                      "full_size": os.path.getsize(p)})
        blob += d
    open(os.path.join(HERE, "corpus_sample.bin"), "wb").write(bytes(blob))

    def ref_outputs(data, chunk):
        n = len(data) // chunk
        rec = {"n": n}
        if n == 0:
            return rec
        for alg, name in O.HASH_NAMES.items():
            rec[name] = O.ref_hash_chunks(alg, data, chunk).tobytes().hex()
        for alg, name in O.COMP_NAMES.items():
            blocks, lens = O.ref_compress_chunks(alg, data, chunk)
            rec[name + "_lens"] = [int(x) for x in lens]
            rec[name + "_sha256"] = [hashlib.sha256(blocks[i, : lens[i]].tobytes()).hexdigest()[:32] for i in range(n)]
        return rec

    ref = {"_source": "oracle/_ref/ref_cli = reference Skein source + vendored liblz4.a 1.8.2 / liblzf.a "
                      "(zeroed table) + OpenSSL SHA256; per-chunk calls as HashAndCompress.cpp:121-134,342-355",
           "files": index, "corpus": {}, "synthetic": {}, "full_corpus_totals": {}}
    blob = np.frombuffer(bytes(blob), np.uint8)
    for ent in index:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        ref["corpus"][ent["name"]] = {str(c): ref_outputs(d, c) for c in CHUNKS}
    # synthetic stream: first 64 chunks at 8 KiB, 32 at 4 KiB, 4 at 64 KiB
    for chunk, n in ((8192, 64), (4096, 32), (65536, 4)):
        d = O.synth(chunk, 0, n)
        rec = ref_outputs(d, chunk)
        rec["data_sha256"] = hashlib.sha256(d.tobytes()).hexdigest()
        ref["synthetic"][str(chunk)] = rec
    # whole-file totals at 8 KiB (SURVEY.md A.4 lists two of them); checked when the full corpus is staged
    for p in files:
        d = np.fromfile(p, np.uint8)
        tot = {}
        for chunk in (4096, 8192):
            n = len(d) // chunk
            if n == 0:
                continue
            t = {"n": n}
            for alg, name in O.COMP_NAMES.items():
                _, lens = O.ref_compress_chunks(alg, d, chunk)
                t[name] = int(lens.sum())
                t[name + "_zero"] = int((lens == 0).sum())
            for alg, name in O.HASH_NAMES.items():
                t[name + "_sha256"] = hashlib.sha256(O.ref_hash_chunks(alg, d, chunk).tobytes()).hexdigest()[:32]
            tot[str(chunk)] = t
        ref["full_corpus_totals"][os.path.relpath(p, O.CORPUS_DIR)] = tot
    with open(os.path.join(HERE, "ref_outputs.json"), "w") as f:
        json.dump(ref, f, separators=(",", ":"))
    print("corpus_sample.bin", len(blob), "bytes;", len(index), "files")


if __name__ == "__main__":
    main()
