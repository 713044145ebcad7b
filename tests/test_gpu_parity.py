"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (compute_war_b200.Engine ->
libcomputewar.so), against (1) the committed golden outputs of the reference's own primitives
(tests/golden/ref_outputs.json), (2) the reference's Skein KAT vectors, and (3) the CPU oracle on seeded
inputs.  Bit-exact everywhere: this path is integer/byte work.  Nothing here reads /root/reference."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

cw = pytest.importorskip("compute_war_b200")


@pytest.fixture(scope="module")
def eng():
    e = cw.Engine(0)
    yield e
    e.close()


def _sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()[:32]


# ------------------------------------------------------------------------------------------------
# hashing
# ------------------------------------------------------------------------------------------------

def test_hash_corpus_sample_matches_reference(eng, golden):
    ref, blob = golden["ref"], golden["blob"]
    checked = 0
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk, rec in ref["corpus"][ent["name"]].items():
            if rec["n"] == 0:
                continue
            for alg, name in O.HASH_NAMES.items():
                got = eng.do_hashing(alg, d, int(chunk))
                assert got.tobytes().hex() == rec[name], (ent["name"], chunk, name)
                checked += rec["n"]
    assert checked > 500


@pytest.mark.parametrize("name", ["ShortMsgKAT_512", "LongMsgKAT_512", "ShortMsgKAT_384", "ShortMsgKAT_256", "LongMsgKAT_256",
                                  "ShortMsgKAT_224"])
def test_skein_kat_vectors(eng, golden, name):
    """every byte-aligned message of the reference's NIST KAT files (reference_code/skein/KAT_MCT/), each hashed as a
    single chunk ON THE GPU with the file's own (state, digest) pair: Skein-512/512, Skein-512/384, Skein-256/256 and
    Skein-256/224 through cw_hash_skein; the 512/512 file also through the cw_hash selector the tools use"""
    k = golden["kat"][name]
    cases = [(b, m, md) for b, m, md in k["cases"] if b % 8 == 0 and b > 0]
    assert len(cases) >= 64
    for bits, msg, md in cases:
        m = np.frombuffer(bytes.fromhex(msg), np.uint8)
        assert eng.hash_skein(k["state"], k["hash"], m, len(m)).tobytes().hex() == md, (name, bits)
        if (k["state"], k["hash"]) == (512, 512):
            assert eng.do_hashing(cw.HASH_SKEIN512, m, len(m)).tobytes().hex() == md, (name, bits)
        if k["state"] == 256:  # the variant the reference's tools run has no KAT of its own: pinned through the oracle
            assert eng.do_hashing(cw.HASH_SKEIN256_128, m, len(m)).tobytes() == O.skein(256, 128, m.tobytes()), (name, bits)


def test_skein_kat_batched_as_chunks(eng, golden):
    """the same vectors as batches: all messages of one length form one call (chunk = that length)"""
    for name in ("ShortMsgKAT_512", "ShortMsgKAT_384", "ShortMsgKAT_256", "ShortMsgKAT_224"):
        k = golden["kat"][name]
        by_len = {}
        for b, m, md in k["cases"]:
            if b % 8 == 0 and b > 0:
                by_len.setdefault(b // 8, []).append((m, md))
        for ln, items in by_len.items():
            d = np.frombuffer(b"".join(bytes.fromhex(m) for m, _ in items), np.uint8)
            got = eng.hash_skein(k["state"], k["hash"], d, ln)
            assert [g.tobytes().hex() for g in got] == [md for _, md in items], (name, ln)


def test_skein_golden_kat(eng, golden):
    """skein_golden_kat.txt: every sequential, un-keyed, byte-aligned case of Skein-256 and Skein-512 whose digest fits
    one output block (digest <= state), on the GPU"""
    n = 0
    for state, hbits, mbits, msg, res in golden["kat"]["golden"]:
        if state in (256, 512) and hbits <= state and hbits % 8 == 0 and mbits % 8 == 0 and mbits > 0:
            m = np.frombuffer(bytes.fromhex(msg), np.uint8)
            assert eng.hash_skein(state, hbits, m, len(m)).tobytes().hex() == res, (state, hbits, mbits)
            n += 1
    assert n >= 60


def test_skein_monte_carlo_512_all_checkpoints(eng, golden):
    """MonteCarlo_512, all 100 checkpoints: 100 x 1000 chained hashes of a 128-byte message (genKAT.c:288-296), every
    one of them a GPU call"""
    k = golden["kat"]["MonteCarlo_512"]
    msg = bytes.fromhex(k["seed"])
    for j in range(100):
        for _ in range(1000):
            md = eng.do_hashing(cw.HASH_SKEIN512, np.frombuffer(msg, np.uint8), 128).tobytes()
            msg = md + msg[:64]
        assert md.hex() == k["md"][j], j


def test_skein_monte_carlo_256_first_checkpoints(eng, golden):
    """MonteCarlo_256 (Skein-256/256), the first 5 checkpoints through cw_hash_skein"""
    k = golden["kat"]["MonteCarlo_256"]
    msg = bytes.fromhex(k["seed"])
    for j in range(5):
        for _ in range(1000):
            md = eng.hash_skein(256, 256, np.frombuffer(msg, np.uint8), 128).tobytes()
            msg = md + msg[:96]
        assert md.hex() == k["md"][j], j


@pytest.mark.parametrize("chunk", [1, 7, 31, 32, 33, 63, 64, 65, 100, 128, 1000, 4096, 4104, 8192, 65536])
def test_hash_ragged_sizes_vs_oracle(eng, chunk):
    rng = np.random.default_rng(chunk)
    n = max(1, min(300, (1 << 20) // chunk))
    d = rng.integers(0, 256, n * chunk, dtype=np.uint8)
    for alg in O.HASH_NAMES:
        assert (eng.do_hashing(alg, d, chunk) == O.hash_chunks(alg, d, chunk)).all(), (alg, chunk)


def test_hash_every_length_up_to_200(eng):
    """every final-block length (a build that mis-hashed 43..51-byte tails was caught only by this sweep)"""
    rng = np.random.default_rng(99)
    for n in range(1, 201):
        d = rng.integers(0, 256, 3 * n, dtype=np.uint8)
        for alg in O.HASH_NAMES:
            assert (eng.do_hashing(alg, d, n) == O.hash_chunks(alg, d, n)).all(), (alg, n)


def test_hash_unaligned_buffers(eng):
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, 64 * 4096 + 64, dtype=np.uint8)
    for off in (1, 3, 8):
        d = raw[off: off + 64 * 4096]
        for alg in O.HASH_NAMES:
            assert (eng.do_hashing(alg, d, 4096) == O.hash_chunks(alg, d, 4096)).all()


def test_empty_batch_and_bad_args(eng):
    assert eng.do_hashing(cw.HASH_SKEIN512, np.zeros(0, np.uint8), 4096).shape == (0, 64)
    b, l = eng.do_compression(cw.COMP_LZ4, np.zeros(100, np.uint8), 4096)  # ragged tail only -> 0 chunks
    assert len(l) == 0
    with pytest.raises(cw.EngineError):
        eng.do_hashing(cw.HASH_SKEIN512, np.zeros(1 << 17, np.uint8), 1 << 17)  # chunk > CW_MAX_CHUNK
    with pytest.raises(cw.EngineError):
        eng.do_compression(cw.COMP_LZ4, np.zeros(8192, np.uint8), 4096, stride=4096)  # stride < bound


# ------------------------------------------------------------------------------------------------
# compression
# ------------------------------------------------------------------------------------------------

def _check_blocks(alg, d, chunk, blocks, lens, rec=None):
    n = len(lens)
    ob, ol = O.compress_chunks(alg, d, chunk)
    assert [int(x) for x in lens] == [int(x) for x in ol], (alg, chunk)
    for i in range(n):
        assert (blocks[i, : lens[i]] == ob[i, : ol[i]]).all(), (alg, chunk, i)
    if rec is not None:
        name = O.COMP_NAMES[alg]
        assert [int(x) for x in lens] == rec[name + "_lens"]
        assert [_sha(blocks[i, : lens[i]]) for i in range(n)] == rec[name + "_sha256"]
    assert O.verify_roundtrip(alg, d, chunk, blocks, lens) == 0


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
def test_compress_corpus_sample_byte_identical_to_reference(eng, golden, alg):
    ref, blob = golden["ref"], golden["blob"]
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk, rec in ref["corpus"][ent["name"]].items():
            if rec["n"] == 0:
                continue
            blocks, lens = eng.do_compression(alg, d, int(chunk), cw.MODE_COMPAT)
            _check_blocks(alg, d, int(chunk), blocks, lens, rec)


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
def test_compress_synthetic_stream_byte_identical(eng, golden, alg):
    for chunk, rec in golden["ref"]["synthetic"].items():
        chunk = int(chunk)
        d = O.synth(chunk, 0, rec["n"])
        blocks, lens = eng.do_compression(alg, d, chunk, cw.MODE_COMPAT)
        _check_blocks(alg, d, chunk, blocks, lens, rec)
    # a longer run of the 8 KiB stream against the oracle
    d = O.synth(8192, 1000, 1024)
    blocks, lens = eng.do_compression(alg, d, 8192, cw.MODE_COMPAT)
    _check_blocks(alg, d, 8192, blocks, lens)


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
def test_compress_edge_cases(eng, alg):
    rng = np.random.default_rng(11)
    cases = []
    for n in (13, 14, 16, 17, 32, 64, 65, 100, 127, 128, 129, 255, 256, 300, 1000, 4096, 8192, 65536):
        cases += [np.zeros(n, np.uint8), np.full(n, 255, np.uint8), rng.integers(0, 256, n, dtype=np.uint8),
                  rng.integers(0, 2, n, dtype=np.uint8), (rng.integers(0, 4, n, dtype=np.uint8) + 65).astype(np.uint8)]
    for p in (1, 2, 3, 4, 5, 7, 8, 13, 30, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 1000):
        cases.append(np.tile(rng.integers(0, 256, p, dtype=np.uint8), 8192 // p + 1)[:8192])
    if alg == O.COMP_LZ4:
        cases += [np.zeros(n, np.uint8) for n in (1, 2, 5, 12)]  # < 13 bytes: one literal run
    runs = np.concatenate([np.full(rng.integers(1, 80), rng.integers(0, 256), np.uint8) for _ in range(400)])[:8192]
    cases.append(runs)
    for c in cases:
        n = len(c)
        blocks, lens = eng.do_compression(alg, c, n, cw.MODE_COMPAT)
        _check_blocks(alg, c, n, blocks, lens)
    # the same inputs in fast mode: byte-equal to the restated fast parse, restored by the restated decoder and (grouped by
    # size, one call each) by the reference's own decoder
    by_size = {}
    for c in cases:
        n = len(c)
        blocks, lens = eng.do_compression(alg, c, n, cw.MODE_FAST)
        fb, fl = O.lz4_fast_chunks(c, n) if alg == O.COMP_LZ4 else O.lzf_fast_chunks(c, n)
        assert int(lens[0]) == int(fl[0]) and (blocks[0, : fl[0]] == fb[0, : fl[0]]).all(), (alg, n)
        assert O.verify_roundtrip(alg, c, n, blocks, lens) == 0, (alg, n)
        if lens[0] == 0:  # LZF: kept raw
            blocks[0, :n] = c
        by_size.setdefault(n, []).append((c, blocks[0].copy(), int(lens[0])))
    if O.have_ref_cli():
        for n, items in by_size.items():
            if n < 16 and alg == O.COMP_LZF:
                continue  # (all raw)
            back = O.ref_decompress_chunks(alg, n, np.stack([b for _, b, _ in items]), np.array([l for _, _, l in items], np.uint32))
            assert (back == np.concatenate([c for c, _, _ in items])).all(), (alg, n)


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
@pytest.mark.parametrize("mode", [0, 1])
def test_compress_ragged_chunk_sizes_and_strides(eng, golden, alg, mode):
    """chunk sizes around every internal switch (staged / in-place reads at 16 KiB, tagged LZF entries at 8 KiB, 16-byte
    alignment of the bulk copy), real data, several chunks per call, odd output strides"""
    blob = golden["blob"]
    for chunk in (3, 13, 100, 1000, 4104, 5000, 8191, 8192, 8193, 8200, 16383, 16384, 16400, 20001, 32768, 50000, 65535, 65536):
        n = min(9, len(blob) // chunk)
        d = blob[7: 7 + n * chunk].copy()  # source offset 7 inside the blob: content differs from the other tests
        stride = cw.compress_bound(alg, chunk) + (5 if chunk % 2 else 0)
        blocks, lens = eng.do_compression(alg, d, chunk, mode, stride=stride)
        if mode == 0:
            _check_blocks(alg, d, chunk, blocks, lens)
        else:
            assert O.verify_roundtrip(alg, d, chunk, blocks, lens) == 0, chunk
            ol = O.compress_chunks(alg, d, chunk)[1]
            tot = sum(int(x) if x else chunk for x in lens)
            assert tot <= 1.02 * sum(int(x) if x else chunk for x in ol) + 16 * n, chunk


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
def test_fast_mode_roundtrips_and_ratio(eng, golden, alg):
    """fast mode: every block must decode back exactly; total size within 1 % of the reference's"""
    ref, blob = golden["ref"], golden["blob"]
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk in (4096, 8192, 65536):
            rec = ref["corpus"][ent["name"]][str(chunk)]
            if rec["n"] == 0:
                continue
            blocks, lens = eng.do_compression(alg, d, chunk, cw.MODE_FAST)
            assert O.verify_roundtrip(alg, d, chunk, blocks, lens) == 0
            if O.have_ref_cli():  # the reference's own LZ4_decompress_safe / lzf_decompress (built from its archives)
                back = O.ref_decompress_chunks(alg, chunk, blocks, lens)
                assert len(back) == rec["n"] * chunk and (back == d[: len(back)]).all(), (ent["name"], chunk)
            # the fast parse is a deterministic function of the chunk: byte-equal to its scalar restatement in the oracle
            fb, fl = O.lz4_fast_chunks(d, chunk) if alg == O.COMP_LZ4 else O.lzf_fast_chunks(d, chunk)
            assert (lens == fl).all() and all((blocks[i, : fl[i]] == fb[i, : fl[i]]).all() for i in range(len(fl))), \
                (ent["name"], chunk)
            mine = sum(int(x) if x else chunk for x in lens)
            theirs = sum(int(x) if x else chunk for x in rec[O.COMP_NAMES[alg] + "_lens"])
            assert mine <= theirs * 1.01, (ent["name"], chunk, mine, theirs)
            # ... and equal to the pinned fast-mode outputs (tests/golden/fast_mode_outputs.json)
            pin = golden["fast"]["corpus"][ent["name"]][str(chunk)]
            name = O.COMP_NAMES[alg]
            assert [int(x) for x in lens] == pin[name + "_lens"], (ent["name"], chunk)
            assert [_sha(blocks[i, : lens[i]]) for i in range(len(lens))] == pin[name + "_sha256"], (ent["name"], chunk)


# ------------------------------------------------------------------------------------------------
# fused pass, async offload, device-pointer API
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("halg,calg", [(O.HASH_SKEIN512, O.COMP_LZ4), (O.HASH_SKEIN256_128, O.COMP_LZ4),
                                       (O.HASH_SKEIN256_128, O.COMP_LZF), (O.HASH_SHA256, O.COMP_LZ4),
                                       (O.HASH_SKEIN512, O.COMP_LZF), (O.HASH_SHA256, O.COMP_LZF)])
@pytest.mark.parametrize("chunk", [4096, 8192, 12288, 16384, 32768])
def test_process_block_fused(eng, golden, halg, calg, chunk):
    """hc_sklz4 / hc_sklzf / hc_shlz4 ... : the fused pass equals the two separate primitives and the reference"""
    ref, blob = golden["ref"], golden["blob"]
    d = np.concatenate([blob[: 40 * 8192], O.synth(chunk, 7, 96 if chunk <= 8192 else 24)])
    dig, blocks, lens = eng.process_block(halg, calg, d, chunk)
    assert (dig == O.hash_chunks(halg, d, chunk)).all()
    _check_blocks(calg, d, chunk, blocks, lens)


@pytest.mark.parametrize("chunk", [8208, 10240, 12288, 16368, 16384])
def test_lz4_compat_chunks_read_in_place_above_8k(eng, golden, chunk):
    """compat LZ4 reads chunks above 8 KiB where they lie (13 parsers per SM instead of 6-8): the stand-alone and the
    fused kernel stay byte-identical to the reference's parse on corpus text and the synthetic stream; a ragged size next
    to it takes the staged kernel"""
    blob = golden["blob"]
    d = np.concatenate([blob[: 48 * 8192], O.synth(chunk, 31, 40)])
    d = d[: len(d) // chunk * chunk]
    blocks, lens = eng.do_compression(O.COMP_LZ4, d, chunk, cw.MODE_COMPAT)
    _check_blocks(O.COMP_LZ4, d, chunk, blocks, lens)
    dig, blocks2, lens2 = eng.process_block(O.HASH_SKEIN512, O.COMP_LZ4, d, chunk)
    assert (dig == O.hash_chunks(O.HASH_SKEIN512, d, chunk)).all()
    assert (lens2 == lens).all()
    mask = np.arange(blocks.shape[1])[None, :] < lens[:, None]
    assert (blocks2[mask] == blocks[mask]).all()
    # a ragged chunk size next to it (not a multiple of 16) is not eligible for the in-place parser: staged kernel, same rule
    rag = chunk + 2
    dr = d[: len(d) // rag * rag]
    blocks3, lens3 = eng.do_compression(O.COMP_LZ4, dr, rag, cw.MODE_COMPAT)
    _check_blocks(O.COMP_LZ4, dr, rag, blocks3, lens3)
    dig3, blocks4, lens4 = eng.process_block(O.HASH_SHA256, O.COMP_LZ4, dr, rag)
    assert (dig3 == O.hash_chunks(O.HASH_SHA256, dr, rag)).all() and (lens4 == lens3).all()


def test_large_batch_spans_pipeline_slices(eng):
    """320 MiB from ordinary memory: 20 staged slices of 16 MiB through the 4 pipeline slots (slots reused five times;
    page-locked buffers, 64 MiB slices: test_pinned_and_pageable_callers_agree); sampled + order-independent checks"""
    chunk, n = 8192, 40960  # 320 MiB
    d = O.synth(chunk, 5000, n)
    dig, blocks, lens = eng.process_block(cw.HASH_SKEIN512, cw.COMP_LZ4, d, chunk)
    idx = np.r_[0:64, 8180:8204, 16370:16400, 32760:32776, n - 64:n]
    sub = np.concatenate([d[i * chunk:(i + 1) * chunk] for i in idx])
    assert (dig[idx] == O.hash_chunks(O.HASH_SKEIN512, sub, chunk)).all()
    ob, ol = O.compress_chunks(O.COMP_LZ4, sub, chunk)
    assert (lens[idx] == ol).all()
    for k, i in enumerate(idx):
        assert (blocks[i, : lens[i]] == ob[k, : ol[k]]).all()
    assert O.verify_roundtrip(O.COMP_LZ4, d, chunk, blocks, lens) == 0  # every one of the 40960 blocks decodes back
    # a second run is identical (no dependence on scheduling order)
    dig2, blocks2, lens2 = eng.process_block(cw.HASH_SKEIN512, cw.COMP_LZ4, d, chunk)
    assert (dig == dig2).all() and (lens == lens2).all()


def test_pinned_and_pageable_callers_agree(eng):
    """host-buffer pass: page-locked caller buffers go to the copy engines as they are (64 MiB slices), ordinary memory is
    staged through pinned slot buffers by the copy helpers (16 MiB slices, results copied out when a slot comes round
    again) -- same digests, lengths and block bytes either way; hash-only and compress-only calls too (LZF: length 0 =
    nothing copied)"""
    chunk, n = 8192, 36864  # 288 MiB: 5 slices of 64 MiB, 18 of 16 MiB
    d = O.synth(chunk, 9000, n)
    dig_a, blocks_a, lens_a = eng.process_block(cw.HASH_SKEIN512, cw.COMP_LZ4, d, chunk)  # pageable
    stride = cw.compress_bound(cw.COMP_LZ4, chunk)
    dp = d.copy()
    out = (np.zeros((n, 64), np.uint8), np.zeros((n, stride), np.uint8), np.zeros(n, np.uint32))
    regs = [dp, *out]
    for a in regs:
        eng.host_register(a)
    try:
        dig_b, blocks_b, lens_b = eng.process_block(cw.HASH_SKEIN512, cw.COMP_LZ4, dp, chunk, out=out)
    finally:
        for a in regs:
            eng.host_unregister(a)
    assert (dig_a == dig_b).all() and (lens_a == lens_b).all()
    mask = np.arange(stride)[None, :] < lens_a[:, None]
    assert (blocks_a[mask] == blocks_b[mask]).all()
    idx = np.r_[0:32, 2040:2056, n - 32:n]
    sub = np.concatenate([d[i * chunk:(i + 1) * chunk] for i in idx])
    assert (dig_a[idx] == O.hash_chunks(O.HASH_SKEIN512, sub, chunk)).all()
    assert O.verify_roundtrip(O.COMP_LZ4, d, chunk, blocks_a, lens_a) == 0
    # hash-only and compress-only through the staged path, several slices each
    m = 6000
    assert (eng.do_hashing(cw.HASH_SHA256, d[: m * chunk], chunk)[idx[:48]] == O.hash_chunks(O.HASH_SHA256, sub[: 48 * chunk], chunk)).all()
    bl, ll = eng.do_compression(cw.COMP_LZF, d[: m * chunk], chunk)
    ob, ol = O.compress_chunks(O.COMP_LZF, d[: 64 * chunk], chunk)
    assert (ll[:64] == ol).all() and all((bl[i, : ol[i]] == ob[i, : ol[i]]).all() for i in range(64))
    assert O.verify_roundtrip(O.COMP_LZF, d[: m * chunk], chunk, bl, ll) == 0


def test_packed_output_pinned_and_pageable_callers_agree(eng):
    """cw_hash_compress_packed: ordinary memory is staged through pinned slot buffers (several slices, slots reused),
    page-locked buffers are copied as they are -- same packed stream, offsets, lengths and digests, and the stream is the
    strided call's blocks back to back"""
    chunk, n = 4096, 40000  # 156 MiB: ten staged slices of 16 MiB, three of 64 MiB
    d = O.synth(chunk, 12000, n)
    dig_a, packed_a, off_a, lens_a = eng.process_block_packed(cw.HASH_SKEIN256_128, cw.COMP_LZ4, d, chunk)  # pageable
    cap = n * cw.compress_bound(cw.COMP_LZ4, chunk)
    dp = d.copy()
    dig_b, packed_b = np.zeros((n, 16), np.uint8), np.zeros(cap, np.uint8)
    off_b, lens_b = np.zeros(n + 1, np.uint64), np.zeros(n, np.uint32)
    regs = [dp, dig_b, packed_b, off_b, lens_b]
    for a in regs:
        eng.host_register(a)
    try:
        eng.process_block_packed_ptr(cw.HASH_SKEIN256_128, cw.COMP_LZ4, cw.MODE_COMPAT, dp.ctypes.data, chunk, n, dig_b.ctypes.data,
                                     packed_b.ctypes.data, cap, off_b.ctypes.data, lens_b.ctypes.data)
    finally:
        for a in regs:
            eng.host_unregister(a)
    total = int(off_a[n])
    assert total == int(lens_a.sum(dtype=np.int64)) == int(off_b[n])
    assert (off_a == off_b).all() and (lens_a == lens_b).all() and (dig_a == dig_b).all()
    assert (off_a[:-1] == np.concatenate([[0], np.cumsum(lens_a[:-1], dtype=np.uint64)])).all()
    assert (packed_a == packed_b[:total]).all()
    _, blocks, lens = eng.process_block(cw.HASH_SKEIN256_128, cw.COMP_LZ4, d, chunk)
    assert (lens == lens_a).all()
    mask = np.arange(blocks.shape[1])[None, :] < lens[:, None]
    assert (blocks[mask] == packed_a).all()  # row-major order of the mask = blocks back to back
    assert (dig_a[:64] == O.hash_chunks(O.HASH_SKEIN256_128, d[: 64 * chunk], chunk)).all()


def test_async_offload_tickets(eng):
    chunk, n = 4096, 2048
    d = O.synth(chunk, 0, n)
    stride = cw.compress_bound(cw.COMP_LZ4, chunk)
    outs, tickets = [], []
    for part in range(4):
        lo, hi = n * part // 4, n * (part + 1) // 4
        dig = np.zeros((hi - lo, 16), np.uint8)
        blocks = np.zeros((hi - lo, stride), np.uint8)
        lens = np.zeros(hi - lo, np.uint32)
        src = d[lo * chunk: hi * chunk]
        outs.append((src, dig, blocks, lens))
        tickets.append(eng.submit(cw.HASH_SKEIN256_128, cw.COMP_LZ4, cw.MODE_COMPAT, src.ctypes.data, chunk, hi - lo,
                                  dig.ctypes.data, blocks.ctypes.data, stride, lens.ctypes.data))
    for t in reversed(tickets):
        eng.wait(t)
    for src, dig, blocks, lens in outs:
        assert (dig == O.hash_chunks(O.HASH_SKEIN256_128, src, chunk)).all()
        _check_blocks(O.COMP_LZ4, src, chunk, blocks, lens)


def test_concurrent_callers_threads_and_contexts(eng):
    """doHashing / doCompression are called from --c-threads worker threads in the reference (HashAndCompress.cpp:398-403):
    several host threads on ONE context, and a second context on the same GPU, all at once"""
    import threading
    chunk, n = 4096, 1500
    eng2 = cw.Engine(0)
    jobs = []
    for t in range(6):
        d = O.synth(chunk, 10000 + 3000 * t, n)
        jobs.append({"eng": eng if t % 3 else eng2, "alg": O.COMP_LZ4 if t % 2 else O.COMP_LZF, "d": d})

    def work(j):
        for _ in range(3):
            j["out"] = j["eng"].process_block(cw.HASH_SKEIN256_128, j["alg"], j["d"], chunk)

    th = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    eng2.close()
    for j in jobs:
        dig, blocks, lens = j["out"]
        assert (dig == O.hash_chunks(O.HASH_SKEIN256_128, j["d"], chunk)).all()
        ob, ol = O.compress_chunks(j["alg"], j["d"], chunk)
        assert (lens == ol).all()
        assert all((blocks[i, : ol[i]] == ob[i, : ol[i]]).all() for i in range(0, n, 7))


def test_small_work_items_from_many_threads_share_batches(eng):
    """the reference's calling pattern (HashAndCompress.cpp:160-183,263-272): 14 worker threads, each handing over one
    32 KiB work item (8 x 4 KiB) at a time.  Results equal the oracle's; the calls were served by far fewer launches;
    other kinds of work interleaved from other threads do not mix"""
    import threading
    chunk, per, items, T = 4096, 8, 120, 14
    data = [O.synth(chunk, 1000 * t, per * items) for t in range(T)]
    out = [None] * T
    b0, c0 = eng.batch_stats()

    def worker(t):
        res = []
        kind = (O.HASH_SKEIN256_128, O.COMP_LZ4, 0) if t % 3 else (O.HASH_SKEIN512, O.COMP_LZF, 0)
        for i in range(items):
            d = data[t][i * per * chunk:(i + 1) * per * chunk]
            res.append(eng.process_block(kind[0], kind[1], d, chunk, kind[2]))
        out[t] = (kind, res)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    b1, c1 = eng.batch_stats()
    assert c1 - c0 == T * items and (b1 - b0) < (c1 - c0)
    for t in range(T):
        (halg, calg, _), res = out[t]
        want_d = O.hash_chunks(halg, data[t], chunk)
        want_b, want_l = O.compress_chunks(calg, data[t], chunk)
        for i, (dig, blocks, lens) in enumerate(res):
            sl = slice(i * per, (i + 1) * per)
            assert (dig == want_d[sl]).all() and (lens == want_l[sl]).all(), (t, i)
            for k in range(per):
                assert (blocks[k, : lens[k]] == want_b[i * per + k, : lens[k]]).all(), (t, i, k)


def test_queued_small_jobs_travel_together(eng):
    """cw_submit: many small jobs queued at once complete (any order), with the oracle's results, in few launches"""
    chunk, per, jobs = 4096, 8, 300
    d = O.synth(chunk, 555, per * jobs)
    stride = cw.compress_bound(O.COMP_LZ4, chunk)
    dig = np.zeros((jobs * per, 16), np.uint8)
    blocks = np.zeros((jobs * per, stride), np.uint8)
    lens = np.zeros(jobs * per, np.uint32)
    b0, c0 = eng.batch_stats()
    tickets = []
    for i in range(jobs):
        o = i * per
        tickets.append(eng.submit(O.HASH_SKEIN256_128, O.COMP_LZ4, 0, d[o * chunk:].ctypes.data, chunk, per, dig[o:].ctypes.data,
                                  blocks[o:].ctypes.data, stride, lens[o:].ctypes.data))
    for tk in reversed(tickets):
        eng.wait(tk)
    b1, c1 = eng.batch_stats()
    assert c1 - c0 == jobs and (b1 - b0) <= jobs // 4
    want_b, want_l = O.compress_chunks(O.COMP_LZ4, d, chunk)
    assert (dig == O.hash_chunks(O.HASH_SKEIN256_128, d, chunk)).all() and (lens == want_l).all()
    assert all((blocks[i, : lens[i]] == want_b[i, : lens[i]]).all() for i in range(jobs * per))


def test_device_pointer_api_and_synth(eng):
    torch = pytest.importorskip("torch")
    chunk, n = 8192, 4096
    dev = torch.device("cuda:0")
    src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
    eng.synth_dev(src.data_ptr(), chunk, 123, n)
    eng.sync()
    host = src.cpu().numpy()
    assert (host == O.synth(chunk, 123, n)).all()  # CUDA generator == CPU generator
    stride = cw.compress_bound(cw.COMP_LZ4, chunk)
    dig = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    blocks = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    before = eng.launch_count()
    eng.hash_compress_dev(cw.HASH_SKEIN512, cw.COMP_LZ4, cw.MODE_COMPAT, src.data_ptr(), chunk, n, dig.data_ptr(),
                          blocks.data_ptr(), stride, lens.data_ptr(), st)
    torch.cuda.synchronize()
    assert eng.launch_count() > before
    dig_h = dig.cpu().numpy().reshape(n, 64)
    lens_h = lens.cpu().numpy().astype(np.uint32)
    blocks_h = blocks.cpu().numpy().reshape(n, stride)
    sel = np.r_[0:128, n - 128:n]
    sub = np.concatenate([host[i * chunk:(i + 1) * chunk] for i in sel])
    assert (dig_h[sel] == O.hash_chunks(O.HASH_SKEIN512, sub, chunk)).all()
    assert O.verify_roundtrip(O.COMP_LZ4, host, chunk, blocks_h, lens_h) == 0
    ob, ol = O.compress_chunks(O.COMP_LZ4, sub, chunk)
    assert (lens_h[sel] == ol).all()


# ------------------------------------------------------------------------------------------------
# decompression: the GPU decoders restore what the reference's compressors produced
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
def test_decompress_reference_streams(eng, golden, alg):
    ref, blob = golden["ref"], golden["blob"]
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk in (4096, 8192, 65536):
            n = len(d) // chunk
            if n == 0:
                continue
            blocks, lens = O.compress_chunks(alg, d, chunk)  # byte-identical to the reference's output (test_oracle)
            if alg == O.COMP_LZF:  # chunks that "did not fit" are stored raw by the caller
                for i in np.nonzero(lens == 0)[0]:
                    blocks[i, :chunk] = d[i * chunk:(i + 1) * chunk]
            out, out_lens = eng.do_decompression(alg, blocks, lens, chunk)
            assert (out_lens == chunk).all(), (ent["name"], chunk)
            assert (out == d[: n * chunk]).all(), (ent["name"], chunk)


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
@pytest.mark.parametrize("mode", [0, 1])
def test_device_roundtrip_every_block(eng, alg, mode):
    """compress then decompress 16384 chunks on the device; every decoded chunk must equal its input"""
    torch = pytest.importorskip("torch")
    chunk, n = 8192, 16384
    dev = torch.device("cuda:0")
    src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
    eng.synth_dev(src.data_ptr(), chunk, 777, n)
    stride = cw.compress_bound(alg, chunk)
    blocks = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    back = torch.zeros(n * chunk, dtype=torch.uint8, device=dev)
    out_lens = torch.empty(n, dtype=torch.int32, device=dev)
    eng.compress_dev(alg, mode, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr())
    eng.sync()
    if alg == O.COMP_LZF:  # store the chunks that did not fit raw, as the reference's caller would
        raw = (lens == 0).nonzero().flatten()
        assert len(raw) > 0
        blocks.view(n, stride)[raw, :chunk] = src.view(n, chunk)[raw]
    torch.cuda.synchronize()
    eng.decompress_dev(alg, blocks.data_ptr(), stride, lens.data_ptr(), chunk, n, back.data_ptr(), out_lens.data_ptr())
    eng.sync()
    assert bool((out_lens == chunk).all())
    assert torch.equal(back, src)


@pytest.mark.parametrize("alg", [O.COMP_LZ4, O.COMP_LZF])
def test_packed_output(eng, alg):
    """cw_pack_dev: offsets = exclusive scan of stored sizes (raw size where LZF returned 0), bytes back to back"""
    torch = pytest.importorskip("torch")
    from compute_war_b200.sharding import packed_offsets
    chunk, n = 8192, 5000  # not a multiple of the 1024-chunk scan tile
    dev = torch.device("cuda:0")
    src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
    eng.synth_dev(src.data_ptr(), chunk, 31, n)
    stride = cw.compress_bound(alg, chunk)
    blocks = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    eng.compress_dev(alg, cw.MODE_COMPAT, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr())
    eng.sync()
    lens_h = lens.cpu().numpy().astype(np.uint32)
    want = packed_offsets(lens_h, chunk)
    packed = torch.zeros(int(want[-1]) + 64, dtype=torch.uint8, device=dev)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    eng.pack_dev(blocks.data_ptr(), stride, lens.data_ptr(), src.data_ptr(), chunk, n, packed.data_ptr(), int(want[-1]),
                 offsets.data_ptr())
    eng.sync()
    assert (offsets.cpu().numpy() == want).all()
    p = packed.cpu().numpy()
    b = blocks.cpu().numpy().reshape(n, stride)
    s_h = src.cpu().numpy().reshape(n, chunk)
    for i in list(range(0, 40)) + list(range(1020, 1030)) + list(range(n - 40, n)):
        seg = p[want[i]: want[i + 1]]
        assert (seg == (b[i, : lens_h[i]] if lens_h[i] else s_h[i])).all(), i
    assert (p[want[-1]:] == 0).all()  # nothing written past the end


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("halg", [O.HASH_SKEIN512, O.HASH_SKEIN256_128, O.HASH_SHA256])
def test_process_block_fused_ragged_sizes(eng, golden, halg, mode):
    """fused pass, both LZ4 modes, chunk sizes whose last hash block has every length 1..64 (and unaligned buffers):
    digests equal the oracle's, blocks decode back"""
    blob = golden["blob"]
    for chunk in list(range(4097, 4097 + 64)) + [13, 100, 8191, 8193, 16400]:
        n = 5
        d = blob[3: 3 + n * chunk].copy()
        dig, blocks, lens = eng.process_block(halg, O.COMP_LZ4, d, chunk, mode)
        assert (dig == O.hash_chunks(halg, d, chunk)).all(), (chunk, mode)
        assert O.verify_roundtrip(O.COMP_LZ4, d, chunk, blocks, lens) == 0, (chunk, mode)


@pytest.mark.parametrize("halg", [O.HASH_SKEIN512, O.HASH_SKEIN256_128, O.HASH_SHA256])
@pytest.mark.parametrize("calg", [O.COMP_LZ4, O.COMP_LZF])
def test_process_block_fused_fast_equals_separate(eng, golden, halg, calg):
    """fast mode of the fused pass (hc_sklz4 / hc_sklzf / hc_shlz4 ...): digests equal the oracle's, blocks equal the restated
    fast parse byte for byte (= what the stand-alone fast kernel writes) and decode back; staged chunks, chunks read in place
    (64 KiB), ragged sizes, incompressible chunks (LZF: length 0)"""
    blob = golden["blob"]
    for chunk in (4096, 8192, 16384, 65536, 5000, 20001):
        n = 24 if chunk <= 8192 else 6
        d = np.concatenate([blob[11: 11 + n * chunk], O.synth(chunk, 3, n)])
        dig, blocks, lens = eng.process_block(halg, calg, d, chunk, cw.MODE_FAST)
        assert (dig == O.hash_chunks(halg, d, chunk)).all(), chunk
        fb, fl = O.lz4_fast_chunks(d, chunk) if calg == O.COMP_LZ4 else O.lzf_fast_chunks(d, chunk)
        assert (lens == fl).all(), (chunk, lens, fl)
        assert all((blocks[i, : fl[i]] == fb[i, : fl[i]]).all() for i in range(len(fl))), chunk
        assert O.verify_roundtrip(calg, d, chunk, blocks, lens) == 0, chunk


@pytest.mark.parametrize("halg,calg,mode", [(O.HASH_SKEIN512, O.COMP_LZ4, 0), (O.HASH_SKEIN512, O.COMP_LZ4, 1),
                                            (O.HASH_SKEIN256_128, O.COMP_LZF, 0), (O.HASH_SKEIN512, O.COMP_LZF, 1),
                                            (O.HASH_SHA256, O.COMP_LZ4, 0)])
def test_packed_host_output(eng, halg, calg, mode):
    """cw_hash_compress_packed: packed bytes / offsets / lens / digests equal the strided call's, across pipeline slices
    (72 MiB > one 64 MiB slice), incompressible chunks included (LZF: stored raw); a short buffer is refused"""
    from compute_war_b200.sharding import packed_offsets
    chunk, n = 8192, 9216 + 17
    d = O.synth(chunk, 5, n)
    dig0, blocks, lens0 = eng.process_block(halg, calg, d, chunk, mode)
    dig, packed, offsets, lens = eng.process_block_packed(halg, calg, d, chunk, mode)
    assert (dig == dig0).all() and (lens == lens0).all()
    want = packed_offsets(lens0, chunk)
    assert (offsets == want).all() and len(packed) == int(want[-1])
    dv = d.reshape(n, chunk)
    for i in list(range(0, 64)) + list(range(8180, 8200)) + list(range(n - 64, n)):  # both sides of the slice boundary
        seg = packed[int(want[i]): int(want[i + 1])]
        assert (seg == (blocks[i, : lens0[i]] if lens0[i] else dv[i])).all(), i
    # all of it: unpack into a strided array and decode on the device
    if calg == O.COMP_LZ4:
        stride = cw.compress_bound(calg, chunk)
        b2 = np.zeros((n, stride), np.uint8)
        for i in range(n):
            b2[i, : lens0[i]] = packed[int(want[i]): int(want[i + 1])]
        assert O.verify_roundtrip(calg, d, chunk, b2, lens0) == 0
    with pytest.raises(cw.EngineError):
        eng.process_block_packed(halg, calg, d, chunk, mode, cap=int(want[-1]) - 1)
    # the context still works after the refusal, and an empty batch yields offsets = [0]
    dig3, packed3, offsets3, _ = eng.process_block_packed(halg, calg, d[: 100 * chunk], chunk, mode)
    assert (dig3 == dig0[:100]).all() and (offsets3 == want[:101]).all()
    assert eng.process_block_packed(halg, calg, d[:0], chunk, mode)[2].tolist() == [0]


def test_pack_grow_then_dedup_on_one_context(eng):
    """dedup, then a pack call that needs more scratch than any before it, then dedup again, on ONE context (ADVICE r01:
    a stray cudaFree in the pack path once freed the dedup table); run under compute-sanitizer by tools/gpu_memcheck.sh"""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    chunk = 4096

    def dedup(n, seed):
        src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
        eng.synth_dev(src.data_ptr(), chunk, seed, n)
        src.view(n, chunk)[n // 2:] = src.view(n, chunk)[: n - n // 2]  # second half repeats the first
        dig = torch.empty(n * 16, dtype=torch.uint8, device=dev)
        eng.hash_dev(cw.HASH_SKEIN256_128, src.data_ptr(), chunk, n, dig.data_ptr())
        first = torch.full((n,), -1, dtype=torch.int32, device=dev)
        uniq = torch.zeros(1, dtype=torch.int64, device=dev)
        eng.dedup_dev(dig.data_ptr(), 16, n, first.data_ptr(), uniq.data_ptr())
        eng.sync()
        want, want_u = O.dedup_first(dig.cpu().numpy().reshape(n, 16))
        assert (first.cpu().numpy().astype(np.uint32) == want).all() and int(uniq.item()) == want_u

    def pack(n):
        src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
        eng.synth_dev(src.data_ptr(), chunk, 3, n)
        stride = cw.compress_bound(O.COMP_LZ4, chunk)
        blocks = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
        lens = torch.empty(n, dtype=torch.int32, device=dev)
        eng.compress_dev(O.COMP_LZ4, cw.MODE_COMPAT, src.data_ptr(), chunk, n, blocks.data_ptr(), stride, lens.data_ptr())
        packed = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        eng.pack_dev(blocks.data_ptr(), stride, lens.data_ptr(), src.data_ptr(), chunk, n, packed.data_ptr(), n * stride,
                     offsets.data_ptr())
        eng.sync()
        assert int(offsets[-1].item()) == int(lens.sum(dtype=torch.int64).item())

    dedup(20000, 1)
    pack(3000)
    pack(60000)  # grows every scratch of the pack path
    dedup(20000, 2)
    dedup(5000, 3)
    pack(100)


# ------------------------------------------------------------------------------------------------
# digest post-processing
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("halg", [O.HASH_SKEIN256_128, O.HASH_SKEIN512, O.HASH_SHA256])
def test_digest_byte_distribution_matches_reference_statistic(eng, halg):
    """cw_digest_hist_dev + cw_digest_byte_stats == Hasher::LogResults' numbers (hashing_randomness/hash.cpp:30-101)"""
    torch = pytest.importorskip("torch")
    chunk, n = 4096, 20000
    D = cw.digest_bytes(halg)
    dev = torch.device("cuda:0")
    src = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
    eng.synth_dev(src.data_ptr(), chunk, 5, n)
    dig = torch.empty(n * D, dtype=torch.uint8, device=dev)
    eng.hash_dev(halg, src.data_ptr(), chunk, n, dig.data_ptr())
    hist = torch.full((D * 256,), -1, dtype=torch.int64, device=dev)
    eng.digest_hist_dev(dig.data_ptr(), D, n, hist.data_ptr())
    eng.sync()
    dig_h = dig.cpu().numpy().reshape(n, D)
    want = np.stack([np.bincount(dig_h[:, b], minlength=256) for b in range(D)]).reshape(-1)
    assert (hist.cpu().numpy() == want).all()
    avg, sd, tot = eng.digest_stats(dig_h, 16)
    o_avg, o_sd, o_tot = O.digest_byte_stats(dig_h, 16)
    assert np.allclose(avg, o_avg, rtol=0, atol=0) and np.allclose(sd, o_sd, rtol=1e-12) and abs(tot - o_tot) <= 1e-12 * o_tot
    # unaligned base, odd width (byte path of the kernel)
    odd = dig[3: 3 + (n - 1) * D]
    hist2 = torch.zeros(7 * 256, dtype=torch.int64, device=dev)
    eng.digest_hist_dev(odd.data_ptr(), 7, (n - 1) * D // 7, hist2.data_ptr())
    eng.sync()
    o = odd.cpu().numpy()[: ((n - 1) * D // 7) * 7].reshape(-1, 7)
    want2 = np.stack([np.bincount(o[:, b], minlength=256) for b in range(7)]).reshape(-1)
    assert (hist2.cpu().numpy() == want2).all()


def test_digest_stats_value_255_quirk(eng):
    """the reference's map has no key 255 until one is seen (hash.cpp:38,55): both branches, tiny inputs"""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 255, size=(300, 16), dtype=np.uint8)  # never 255
    b = a.copy()
    b[::7, 3] = 255
    for d in (a, b, a[:1], np.zeros((0, 16), np.uint8)):
        avg, sd, tot = eng.digest_stats(d, 16)
        o_avg, o_sd, o_tot = O.digest_byte_stats(d, 16)
        assert (avg == o_avg).all() and np.allclose(sd, o_sd, rtol=1e-12, atol=0) and abs(tot - o_tot) <= 1e-12 * max(o_tot, 1)


@pytest.mark.parametrize("halg", [O.HASH_SKEIN256_128, O.HASH_SKEIN512])
def test_dedup_map(eng, halg):
    """first[i] = first chunk with the same digest; duplicates planted by repeating chunks"""
    torch = pytest.importorskip("torch")
    chunk, n = 4096, 30000
    D = cw.digest_bytes(halg)
    dev = torch.device("cuda:0")
    base = torch.empty(n * chunk, dtype=torch.uint8, device=dev)
    eng.synth_dev(base.data_ptr(), chunk, 77, n)
    rng = np.random.default_rng(11)
    pick = np.arange(n)
    dup = rng.random(n) < 0.4
    pick[dup] = rng.integers(0, 500, size=int(dup.sum()))  # 40 % of the chunks repeat one of 500 chunks
    src = base.view(n, chunk)[torch.from_numpy(pick).to(dev)].contiguous()
    dig = torch.empty(n * D, dtype=torch.uint8, device=dev)
    eng.hash_dev(halg, src.data_ptr(), chunk, n, dig.data_ptr())
    first = torch.full((n,), -1, dtype=torch.int32, device=dev)
    uniq = torch.zeros(1, dtype=torch.int64, device=dev)
    eng.dedup_dev(dig.data_ptr(), D, n, first.data_ptr(), uniq.data_ptr())
    eng.sync()
    want, want_u = O.dedup_first(dig.cpu().numpy().reshape(n, D))
    assert (first.cpu().numpy().astype(np.uint32) == want).all()
    assert int(uniq.item()) == want_u and want_u < n
    # structured "digests" (all equal / two values) still resolve
    z = torch.zeros(1000 * 16, dtype=torch.uint8, device=dev)
    z[16 * 500:] = 9
    f2 = torch.empty(1000, dtype=torch.int32, device=dev)
    eng.dedup_dev(z.data_ptr(), 16, 1000, f2.data_ptr(), uniq.data_ptr())
    eng.sync()
    f2h = f2.cpu().numpy()
    assert (f2h[:500] == 0).all() and (f2h[500:] == 500).all() and int(uniq.item()) == 2


def test_host_tool_end_to_end(eng, tmp_path):
    """the C++ host tool (reference CLI, C ABI underneath): result line, dumped outputs == oracle, randomness lines,
    --gpu-offload batches, and -- when the box has them -- two GPUs give the same bytes as one"""
    import os
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compute_war_b200", "hashandcompress")
    if not os.path.exists(exe):
        pytest.skip("host tool not built")
    chunk, n = 4096, 3000
    d = O.synth(chunk, 900, n)
    f = tmp_path / "in.bin"
    f.write_bytes(d.tobytes() + b"tail that is dropped")
    want_dig = O.hash_chunks(O.HASH_SKEIN256_128, d, chunk)
    ob, ol = O.compress_chunks(O.COMP_LZF, d, chunk)
    want_blocks = b"".join(bytes(ob[i, : ol[i]]) if ol[i] else bytes(d[i * chunk:(i + 1) * chunk]) for i in range(n))

    def run(*extra):
        pre = str(tmp_path / ("o" + "_".join(extra).replace("-", "")))
        r = subprocess.run([exe, "-H", "skein", "-C", "lzf", "-k", str(chunk), "--dump", pre, *extra, str(f)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert re.match(r"^skein\|lzf\|\d+\|\d+$", r.stdout.splitlines()[0]), r.stdout
        assert np.fromfile(pre + ".digests", np.uint8).tobytes() == want_dig.tobytes()
        assert (np.fromfile(pre + ".lens", np.uint32) == ol).all()
        assert open(pre + ".blocks", "rb").read() == want_blocks
        return r.stdout

    out = run("--randomness")
    lines = out.splitlines()[1:]
    o_avg, o_sd, o_tot = O.digest_byte_stats(want_dig, 16)
    for b in range(16):
        k, alg, avg, sd = lines[b].split("|")[:4]
        assert int(k) == b and alg == "skein" and abs(float(avg) - o_avg[b]) < 1e-3 * o_avg[b] and abs(float(sd) - o_sd[b]) < 1e-3 * o_sd[b]
    assert lines[16].startswith("Average std dev: ") and abs(float(lines[16].split(": ")[1]) - o_tot) < 1e-3 * o_tot
    run("--gpu-offload", "1", "--c-threads", "5")
    run("--no-pin")  # pageable host buffers, like the reference's malloc'd ones
    # the reference's own calling pattern: 14 worker threads, one call per work item of 8 blocks
    r = subprocess.run([exe, "-H", "skein", "-C", "lzf", "-k", str(chunk), "--work-items", "-c", "14", "-r", "8", "--no-pin",
                        "--dump", str(tmp_path / "wi"), str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert np.fromfile(str(tmp_path / "wi.digests"), np.uint8).tobytes() == want_dig.tobytes()
    assert open(str(tmp_path / "wi.blocks"), "rb").read() == want_blocks
    m = re.search(r"(\d+) calls served by (\d+) launches", r.stderr)
    assert m and int(m.group(1)) >= n // 8 and int(m.group(2)) < int(m.group(1))
    # the same items, HashOffload style: started asynchronously, many in flight per worker
    r = subprocess.run([exe, "-H", "skein", "-C", "lzf", "-k", str(chunk), "--work-items", "--gpu-offload", "1", "-c", "14", "-r", "8",
                        "--no-pin", "--dump", str(tmp_path / "wo"), str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert np.fromfile(str(tmp_path / "wo.digests"), np.uint8).tobytes() == want_dig.tobytes()
    assert open(str(tmp_path / "wo.blocks"), "rb").read() == want_blocks
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() >= 2:
        run("--devices", "2")


def test_host_tool_prints_the_reference_tools_log_lines(eng, tmp_path):
    """--experiment / --hash-perf: the per-block lines of the reference's `experiment` and hashing_perf tools
    (experiment.cpp:114,122,253,260; hash.cpp:18-23,38-43), read the way its notebooks read them
    (notebooks/compression-comparison.ipynb cell 1, hash-perf.ipynb cells 3-4)"""
    import os
    import subprocess
    pd = pytest.importorskip("pandas")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compute_war_b200", "hashandcompress")
    if not os.path.exists(exe):
        pytest.skip("host tool not built")
    chunk, n = 4096, 500
    d = O.synth(chunk, 4242, n)
    f = tmp_path / "blocks.bin"
    f.write_bytes(d.tobytes() + b"partial block")
    r = subprocess.run([exe, "--experiment", "-C", "both", str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    log = tmp_path / "exp.out"
    log.write_text(r.stdout)
    df = pd.read_csv(str(log), sep="|", header=None)
    df.columns = ["alg", "comp_size", "compress_us", "decompress_us", "file", "block"]
    assert len(df) == 2 * n and list(df.alg[:2]) == ["lzf", "lz4"] and int(df.block.min()) == 1 and int(df.block.max()) == n
    for name, alg in (("lz4", O.COMP_LZ4), ("lzf", O.COMP_LZF)):
        sub = df[df.alg == name].sort_values("block")
        assert (sub.comp_size.to_numpy() == O.compress_chunks(alg, d, chunk)[1]).all()  # lzf: 0 where it did not fit, as the reference prints
        assert (sub.compress_us > 0).all() and (sub.decompress_us > 0).all() and (sub.file == str(f)).all()
    r = subprocess.run([exe, "--hash-perf", str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    log = tmp_path / "hash_perf_sb.1"
    log.write_text(r.stdout)
    sb = pd.read_csv(str(log), sep="|", header=None).drop(4, axis=1)
    sb.columns = "FilePath|BlockIndex|Algo|TimeElapsedUS".split("|")
    assert len(sb) == 2 * n and list(sb.BlockIndex) == list(range(2 * n))
    g = sb.groupby("Algo").TimeElapsedUS.mean()
    assert g["Skein256"] > 0 and g["Sha256"] > 0 and set(sb.Algo) == {"Skein256", "Sha256"}


def test_decompress_rejects_malformed_blocks(eng):
    rng = np.random.default_rng(3)
    chunk, n = 4096, 64
    stride = cw.compress_bound(cw.COMP_LZ4, chunk)
    junk = rng.integers(0, 256, (n, stride), dtype=np.uint8)
    lens = rng.integers(1, stride, n).astype(np.uint32)
    out, out_lens = eng.do_decompression(cw.COMP_LZ4, junk, lens, chunk)
    assert (out_lens == 0).sum() >= n - 1  # random bytes are (almost) never a valid 4096-byte LZ4 block
    good_b, good_l = O.compress_chunks(O.COMP_LZ4, np.zeros(chunk, np.uint8), chunk)
    trunc = good_l.copy()
    trunc[0] -= 1  # truncated stream
    _, ol = eng.do_decompression(cw.COMP_LZ4, good_b, trunc, chunk)
    assert ol[0] == 0


def test_full_corpus_totals_if_staged(eng, golden):
    """whole Canterbury + Canterbury-large at 4 KiB and 8 KiB against the totals recorded from the reference"""
    import os
    files = O.corpus_files()
    if not files:
        pytest.skip("full corpus not staged (oracle/_ref/corpus)")
    for p in files:
        d = np.fromfile(p, np.uint8)
        for chunk in (4096, 8192):
            tot = golden["ref"]["full_corpus_totals"][os.path.relpath(p, O.CORPUS_DIR)].get(str(chunk))
            if not tot:
                continue
            for alg, name in O.COMP_NAMES.items():
                _, lens = eng.do_compression(alg, d, chunk)
                assert int(lens.sum()) == tot[name] and int((lens == 0).sum()) == tot[name + "_zero"], (p, chunk, name)
            for alg, name in O.HASH_NAMES.items():
                dig = eng.do_hashing(alg, d, chunk)
                assert _sha(dig.tobytes()) == tot[name + "_sha256"], (p, chunk, name)
