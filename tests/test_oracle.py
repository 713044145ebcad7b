"""CPU tests: pin the oracle (oracle/cw_oracle.c) against the reference's golden vectors
(tests/golden/skein_kat.json <- reference_code/skein/KAT_MCT) and against outputs of the reference's own
primitives (tests/golden/ref_outputs.json <- oracle/_ref/ref_cli).  No GPU, no /root/reference needed."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O


@pytest.mark.parametrize("name", ["ShortMsgKAT_512", "ShortMsgKAT_384", "ShortMsgKAT_256", "ShortMsgKAT_224",
                                  "LongMsgKAT_512", "LongMsgKAT_256"])
def test_skein_nist_kat(golden, name):
    k = golden["kat"][name]
    assert len(k["cases"]) >= 65
    for bits, msg, md in k["cases"]:
        got = O.skein(k["state"], k["hash"], bytes.fromhex(msg), bits)
        assert got.hex() == md, (name, bits)


def test_skein_golden_kat(golden):
    cases = golden["kat"]["golden"]
    assert len(cases) == 228
    for state, hbits, mbits, msg, res in cases:
        assert O.skein(state, hbits, bytes.fromhex(msg), mbits).hex() == res, (state, hbits, mbits)


def test_skein_known_answer_ff():
    # skein_golden_kat.txt:2434-2442
    assert O.skein(512, 512, b"\xff").hex().startswith("8fca8d2705f99a56904308a4004c64ef")


@pytest.mark.parametrize("name", ["MonteCarlo_512", "MonteCarlo_256"])
def test_skein_monte_carlo(golden, name):
    k = golden["kat"][name]
    msg = bytes.fromhex(k["seed"])
    nb = k["hash"] // 8
    for j in range(3 if name.endswith("256") else 5):  # each checkpoint = 1000 chained hashes (genKAT.c:288-296)
        for _ in range(1000):
            md = O.skein(k["state"], k["hash"], msg)
            msg = md + msg[: 128 - nb]
        assert md.hex() == k["md"][j], j


@pytest.mark.slow
def test_skein_extremely_long(golden):
    k = golden["kat"]["ExtremelyLong_512"]
    reps = 16777216 // 64  # 16 MiB prefix is not the KAT; the full 1 GiB one takes ~10 s, run it whole
    msg = np.frombuffer((k["text"].encode() * 4096), np.uint8)
    big = np.tile(msg, 16777216 // 4096)
    assert reps > 0 and big.size == 1 << 30
    assert O.skein(512, 512, big).hex() == k["md"]


def test_sha256_known_answers():
    assert O.sha256(b"").hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert O.sha256(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    m = b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"
    assert O.sha256(m).hex() == "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"
    rng = np.random.default_rng(1)
    for n in list(range(0, 130)) + [4095, 4096, 4097, 8192, 65536]:
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert O.sha256(b) == hashlib.sha256(b).digest(), n


def _check_against_ref(data, chunk, rec):
    n = rec["n"]
    assert n == len(data) // chunk
    if n == 0:
        return
    for alg, name in O.HASH_NAMES.items():
        assert O.hash_chunks(alg, data, chunk).tobytes().hex() == rec[name], (name, chunk)
    for alg, name in O.COMP_NAMES.items():
        blocks, lens = O.compress_chunks(alg, data, chunk)
        assert [int(x) for x in lens] == rec[name + "_lens"], (name, chunk)
        for i in range(n):
            assert hashlib.sha256(blocks[i, : lens[i]].tobytes()).hexdigest()[:32] == rec[name + "_sha256"][i]
        assert O.verify_roundtrip(alg, data, chunk, blocks, lens) == 0


def test_oracle_matches_reference_on_corpus_sample(golden):
    ref, blob = golden["ref"], golden["blob"]
    assert len(ref["files"]) == 13
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk, rec in ref["corpus"][ent["name"]].items():
            _check_against_ref(d, int(chunk), rec)


def test_oracle_matches_reference_on_synthetic(golden):
    for chunk, rec in golden["ref"]["synthetic"].items():
        chunk = int(chunk)
        d = O.synth(chunk, 0, rec["n"])
        assert hashlib.sha256(d.tobytes()).hexdigest() == rec["data_sha256"]  # pins the generator itself
        _check_against_ref(d, chunk, rec)


def test_smoke_values_from_survey(golden):
    ent = next(e for e in golden["ref"]["files"] if e["name"].endswith("alice29.txt"))
    d = golden["blob"][ent["offset"]: ent["offset"] + ent["size"]]
    assert O.skein(256, 128, d[:4096]).hex() == "4159727875aebb36b49ca87786a9ddf8"
    assert O.skein(512, 512, d[:8192]).hex().startswith("203aafcd72696bcc2cd22968bd39235e")
    assert int(O.compress_chunks(O.COMP_LZ4, d[:4096], 4096)[1][0]) == 3044
    assert int(O.compress_chunks(O.COMP_LZF, d[:4096], 4096)[1][0]) == 2809


def test_lz4_edge_cases_roundtrip():
    rng = np.random.default_rng(7)
    cases = [np.zeros(n, np.uint8) for n in (1, 4, 12, 13, 14, 16, 17, 64, 300, 4096, 65536)]
    cases += [np.full(n, 0xFF, np.uint8) for n in (13, 4096)]
    cases += [rng.integers(0, 256, n, dtype=np.uint8) for n in (1, 12, 13, 100, 4096, 8192, 65536)]
    cases += [np.tile(rng.integers(0, 256, p, dtype=np.uint8), 8192 // p + 1)[:8192] for p in (1, 2, 3, 5, 30, 255, 256, 257)]
    cases += [np.frombuffer(b"abcd" * 4, np.uint8), np.frombuffer(b"a" * 13, np.uint8)]
    for c in cases:
        n = len(c)
        for alg in (O.COMP_LZ4, O.COMP_LZF):
            if alg == O.COMP_LZF and n < 2:
                continue
            blocks, lens = O.compress_chunks(alg, c, n)
            assert O.verify_roundtrip(alg, c, n, blocks, lens) == 0
            if alg == O.COMP_LZ4:
                assert 0 < lens[0] <= O.compress_bound(alg, n)
    # known streams: 13 x 'a' and "abcd"x4
    b, l = O.compress_chunks(O.COMP_LZ4, np.frombuffer(b"a" * 13, np.uint8), 13)
    assert b[0, : l[0]].tobytes().hex() == "13610100506161616161"
    # incompressible chunk: LZF returns 0 ("did not fit in chunk-1"), LZ4 expands
    r = rng.integers(0, 256, 8192, dtype=np.uint8)
    assert O.compress_chunks(O.COMP_LZF, r, 8192)[1][0] == 0
    assert O.compress_chunks(O.COMP_LZ4, r, 8192)[1][0] > 8192


@pytest.mark.skipif(not (O.have_ref_cli() and O.have_skref()), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_reference_artifacts_full_corpus(golden):
    """whole staged corpus at 8 KiB against the per-file totals recorded from the reference"""
    sk = O.SkRef()
    files = O.corpus_files()
    if not files:
        pytest.skip("corpus not staged")
    import os
    for p in files:
        d = np.fromfile(p, np.uint8)
        tot = golden["ref"]["full_corpus_totals"][os.path.relpath(p, O.CORPUS_DIR)].get("8192")
        if not tot:
            continue
        for alg, name in O.COMP_NAMES.items():
            _, lens = O.compress_chunks(alg, d, 8192)
            assert int(lens.sum()) == tot[name] and int((lens == 0).sum()) == tot[name + "_zero"]
        for alg, name in O.HASH_NAMES.items():
            dig = O.hash_chunks(alg, d, 8192)
            assert hashlib.sha256(dig.tobytes()).hexdigest()[:32] == tot[name + "_sha256"]
        assert sk.skein512(d[:8192].tobytes()) == O.skein(512, 512, d[:8192].tobytes())
        assert sk.skein256(d[:4096].tobytes()) == O.skein(256, 128, d[:4096].tobytes())


def test_digest_statistic_restatement():
    """cwo_digest_byte_stats follows hashing_randomness/hash.cpp:30-101, including its map that lacks the key 255
    until one is seen: checked against the same formula written with a dict, as the reference does it"""
    import math
    rng = np.random.default_rng(5)
    for with255 in (False, True):
        d = rng.integers(0, 255, size=(777, 16), dtype=np.uint8)
        if with255:
            d[::5, 2] = 255
        avg, sd, tot = O.digest_byte_stats(d, 16)
        acc = 0.0
        for b in range(16):
            m = {v: 0 for v in range(255)}
            for v in d[:, b]:
                m[int(v)] = m.get(int(v), 0) + 1
            total = sum(m.values())
            a = total / 255
            s2 = math.sqrt(sum((c - a) ** 2 for _, c in sorted(m.items())) / 254)
            assert avg[b] == a and math.isclose(sd[b], s2, rel_tol=1e-13)
            acc += s2
        assert math.isclose(tot, acc / 16, rel_tol=1e-13)


def test_dedup_definition():
    d = np.array([[1, 2], [3, 4], [1, 2], [0, 0], [3, 4], [1, 2]], dtype=np.uint8)
    first, uniq = O.dedup_first(d)
    assert first.tolist() == [0, 1, 0, 3, 1, 0] and uniq == 3


def test_fast_mode_restatement_matches_pinned_outputs(golden):
    """the restated fast parses (LZ4 and LZF) still produce the blocks recorded in tests/golden/fast_mode_outputs.json"""
    import hashlib
    ref, blob, fast = golden["ref"], golden["blob"], golden["fast"]
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk, rec in fast["corpus"][ent["name"]].items():
            chunk = int(chunk)
            for name, fn in (("lz4", O.lz4_fast_chunks), ("lzf", O.lzf_fast_chunks)):
                blocks, lens = fn(d, chunk)
                assert [int(x) for x in lens] == rec[name + "_lens"], (ent["name"], chunk, name)
                assert [hashlib.sha256(blocks[i, : lens[i]].tobytes()).hexdigest()[:32] for i in range(rec["n"])] == rec[name + "_sha256"]
