"""CPU tests of the product's DEVICE source: csrc/skein.cuh + csrc/sha256.cuh compiled as plain C++, and the
warp-cooperative compressors (csrc/lz4_warp.cuh, csrc/lzf_warp.cuh) run on the lane-accurate warp emulator of
tests/simt/, all compared bit-for-bit with the oracle.  This is how the kernels' algorithms are checked in the
build container, which has no GPU; the same comparisons run on the real GPU under -m gpu."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
CSRC = os.path.join(os.path.dirname(HERE), "compute_war_b200", "csrc")


def _build(name, sources, extra=()):
    so = os.path.join(SIMT, name)
    deps = [os.path.join(SIMT, s) for s in sources] + [os.path.join(SIMT, "warp_emu.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", SIMT, *extra, "-o", so,
                               *[os.path.join(SIMT, s) for s in sources]])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def hosthash():
    return _build("libhosthash.so", ["host_hash.cpp"])


@pytest.fixture(scope="module")
def emu():
    L = _build("libemu.so", ["emu_codecs.cpp", "warp_emu.cpp"])
    L.emu_lz4_compress.restype = C.c_uint32
    if hasattr(L, "emu_lzf_compress"):
        L.emu_lzf_compress.restype = C.c_uint32
        L.emu_lzf_compress_mode.restype = C.c_uint32
    return L


def test_device_hash_source_on_host(hosthash):
    rng = np.random.default_rng(0)
    for n in [0, 1, 7, 8, 31, 32, 33, 55, 56, 63, 64, 65, 127, 128, 129, 200, 1000, 4096, 8192, 65536]:
        b = rng.integers(0, 256, n + 1, dtype=np.uint8)
        bb = b[:n].tobytes()
        for unroll in (0, 1, 2, 3, 4, 5):  # both round-loop shapes, without and with the next-block prefetch loop
            out = np.zeros(64, np.uint8)
            hosthash.host_skein512(b.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), unroll)
            assert out.tobytes() == O.skein(512, 512, bb), (n, unroll)
        for prefetch in (0, 1, 2):
            out = np.zeros(16, np.uint8)
            hosthash.host_skein256_128(b.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), prefetch)
            assert out.tobytes() == O.skein(256, 128, bb), (n, prefetch)
        out = np.zeros(32, np.uint8)
        hosthash.host_sha256(b.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
        assert out.tobytes() == O.sha256(bb), n


def test_short_rings(emu, golden):
    """the kernels that are short of shared memory run parser -> emitter rings of 32 records instead of 64"""
    emu.emu_set_qlen(32)
    try:
        blob = golden["blob"]
        for c in [blob[:8192].copy(), blob[20000:24096].copy(), O.synth(8192, 3, 1), blob[:65536].copy()]:
            n = len(c)
            for alg, fn in ((O.COMP_LZ4, None), (O.COMP_LZF, None)):
                blocks, lens = O.compress_chunks(alg, c, n)
                out = np.zeros(n + n // 255 + 64, np.uint8)
                if alg == O.COMP_LZ4:
                    r = emu.emu_lz4_compress(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
                else:
                    r = emu.emu_lzf_compress(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), n - 1)
                assert r == lens[0] and (out[:r] == blocks[0, :r]).all(), (alg, n)
    finally:
        emu.emu_set_qlen(64)


def _cases(golden, small):
    rng = np.random.default_rng(3)
    cases = []
    blob = golden["blob"]
    for ent in golden["ref"]["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk, cnt in ((4096, 1), (8192, 1), (100, 2), (13, 1), (1000, 1)):
            n = len(d) // chunk
            for i in rng.choice(n, min(cnt, n), replace=False) if n else []:
                cases.append(d[i * chunk:(i + 1) * chunk].copy())
    for n in (13, 14, 16, 17, 31, 32, 33, 64, 65, 66, 67, 100, 127, 128, 129, 255, 256, 300, 1000, 4096):
        cases += [np.zeros(n, np.uint8), rng.integers(0, 256, n, dtype=np.uint8), rng.integers(0, 2, n, dtype=np.uint8)]
    for p in (1, 2, 3, 4, 7, 8, 30, 32, 33, 64, 255, 256, 257):
        cases.append(np.tile(rng.integers(0, 256, p, dtype=np.uint8), 4096 // p + 1)[:4096])
    for i in range(12):
        cases.append(O.synth(8192, i, 1))
    cases += [O.synth(65536, 1, 1), rng.integers(0, 256, 65536, dtype=np.uint8), np.zeros(65536, np.uint8)]
    big = [e for e in golden["ref"]["files"] if e["size"] >= 65536][:3]  # 64 KB of real text: matches that run to the end
    cases += [blob[e["offset"]: e["offset"] + 65536].copy() for e in big]
    if small:
        cases += [np.zeros(n, np.uint8) for n in (1, 2, 5, 12)]
    return cases


@pytest.mark.parametrize("chain", [0, 1])
def test_lz4_warp_algorithm_byte_identical(emu, golden, chain):
    """chain = 1: the build-time option CW_LZ4_CHAIN (batched replay of twin-free windows; off in the shipped library)"""
    L = emu
    if chain:
        L = _build("libemu_chain.so", ["emu_codecs.cpp", "warp_emu.cpp"], extra=("-DCW_LZ4_CHAIN=1",))
        L.emu_lz4_compress.restype = C.c_uint32
    for c in _cases(golden, small=True):
        n = len(c)
        out = np.zeros(n + n // 255 + 64, np.uint8)
        r = L.emu_lz4_compress(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
        blocks, lens = O.compress_chunks(O.COMP_LZ4, c, n)
        assert r == lens[0] and (out[:r] == blocks[0, :r]).all(), n


def test_lz4_direct_mode_and_fast_table(emu, golden):
    """DIRECT mode (chunk read in place, guard page behind it) is byte-identical; the 4096-entry table of fast mode
    yields a valid stream within 1 % of the reference's size at 8 KiB"""
    emu.emu_lz4_compress_direct.restype = C.c_uint32
    emu.emu_lz4_compress_hl.restype = C.c_uint32
    rng = np.random.default_rng(4)
    blob = golden["blob"]
    cases = [blob[:65536].copy(), blob[100000:100000 + 32768].copy(), O.synth(65536, 2, 1), np.zeros(65536, np.uint8),
             rng.integers(0, 256, 65536, dtype=np.uint8), blob[4096:4096 + 16].copy(), blob[0:64].copy()]
    for c in cases:
        n = len(c)
        out = np.zeros(n + n // 255 + 64, np.uint8)
        r = emu.emu_lz4_compress_direct(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), 13)
        blocks, lens = O.compress_chunks(O.COMP_LZ4, c, n)
        assert r == lens[0] and (out[:r] == blocks[0, :r]).all(), n
    tot_fast = tot_ref = 0
    for ent in golden["ref"]["files"][:6]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for i in range(min(3, len(d) // 8192)):
            c = d[i * 8192:(i + 1) * 8192].copy()
            out = np.zeros(8192 + 64, np.uint8)
            r = emu.emu_lz4_compress_hl(c.ctypes.data_as(C.c_void_p), 8192, out.ctypes.data_as(C.c_void_p), 12)
            assert O.verify_roundtrip(O.COMP_LZ4, c, 8192, out[None, :], np.array([r], np.uint32)) == 0
            tot_fast += r
            tot_ref += int(O.compress_chunks(O.COMP_LZ4, c, 8192)[1][0])
    assert tot_fast <= 1.01 * tot_ref


@pytest.mark.parametrize("lowest_lane_wins", [1, 0])
def test_lz4_fast_parse_on_emulator(emu, golden, lowest_lane_wins):
    """fast mode (csrc/lz4_fast.cuh on the warp emulator): byte-identical to its scalar restatement in the oracle whichever
    lane wins a same-address store race, every block restored by the reference's LZ4_decompress_safe (oracle/_ref/ref_cli
    when it is built, the restated decoder otherwise), every table size, staged and in-place (DIRECT) source"""
    emu.emu_lz4_fast.restype = C.c_uint32
    emu.emu_set_store_winner(lowest_lane_wins)
    try:
        cases = _cases(golden, small=True)
        if not lowest_lane_wins:
            cases = cases[::3]
        for c in cases:
            n = len(c)
            for hl, direct in ((0, 0), (13, 0), (14, 1 if n % 4 == 0 else 0)):
                blocks, lens = O.lz4_fast_chunks(c, n, hashlog=hl) if n else (None, [0])
                # one to four warps per chunk (the kernels run teams of 2-6)
                for nteam in ((1, 2, 4) if lowest_lane_wins else (3,)):
                    out = np.zeros(n + n // 255 + 64, np.uint8)
                    r = emu.emu_lz4_fast(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), hl, direct, nteam)
                    assert r == lens[0] and (out[:r] == blocks[0, :r]).all(), (n, hl, direct, nteam)
                assert O.verify_roundtrip(O.COMP_LZ4, c, n, out[None, :], np.array([r], np.uint32)) == 0, (n, hl)
    finally:
        emu.emu_set_store_winner(0)


def test_lz4_fast_restatement_ratio_and_reference_decoder(golden):
    """the restated fast parse over the golden corpus sample: every block decodes with the reference's decoder and every
    file stays within 1 % of the reference's compressed size at 4 K / 8 K / 64 K (the bound north_star sets)"""
    ref, blob = golden["ref"], golden["blob"]
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk in (4096, 8192, 65536):
            rec = ref["corpus"][ent["name"]][str(chunk)]
            if rec["n"] == 0:
                continue
            blocks, lens = O.lz4_fast_chunks(d, chunk)
            if O.have_ref_cli():
                back = O.ref_decompress_chunks(O.COMP_LZ4, chunk, blocks, lens)
                assert (back == d[: len(back)]).all() and len(back) == rec["n"] * chunk
            else:
                assert O.verify_roundtrip(O.COMP_LZ4, d, chunk, blocks, lens) == 0
            assert int(lens.sum()) <= 1.01 * sum(rec["lz4_lens"]), (ent["name"], chunk)


def _raw_where_zero(d, chunk, blocks, lens):
    """LZF convention: length 0 = the chunk did not shrink and is kept raw (lzf.h:59-64); the decoders take it from the block"""
    b = blocks.copy()
    for i in np.nonzero(np.asarray(lens) == 0)[0]:
        b[i, :chunk] = d[i * chunk: (i + 1) * chunk]
    return b


@pytest.mark.parametrize("lowest_lane_wins", [1, 0])
def test_lzf_fast_parse_on_emulator(emu, golden, lowest_lane_wins):
    """LZF fast mode (the LZF flavour of csrc/lz4_fast.cuh on the warp emulator): byte-identical to its scalar restatement in
    the oracle whichever lane wins a store race, both table sizes, staged and in-place source, teams of 1-6 warps; blocks
    restored by the restated lzf decoder (the reference's own decoder runs in the test below)"""
    emu.emu_lzf_fast.restype = C.c_uint32
    emu.emu_set_store_winner(lowest_lane_wins)
    try:
        cases = _cases(golden, small=True)
        if not lowest_lane_wins:
            cases = cases[::3]
        for c in cases:
            n = len(c)
            for hl, direct in ((0, 0), (13, 0), (14, 1 if n % 4 == 0 else 0)):
                blocks, lens = O.lzf_fast_chunks(c, n, hashlog=hl) if n else (None, [0])
                for nteam in ((1, 3, 6) if lowest_lane_wins else (2,)):
                    out = np.zeros(n + n // 32 + 64, np.uint8)
                    r = emu.emu_lzf_fast(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), hl, direct, nteam)
                    assert r == lens[0] and (out[:r] == blocks[0, :r]).all(), (n, hl, direct, nteam)
                if n and r:
                    assert O.verify_roundtrip(O.COMP_LZF, c, n, out[None, :], np.array([r], np.uint32)) == 0, (n, hl)
    finally:
        emu.emu_set_store_winner(0)


def test_lzf_fast_restatement_ratio_and_reference_decoder(golden):
    """the restated LZF fast parse over the golden corpus sample: every block decodes with the reference's lzf_decompress and
    every file stays within 1 % of the reference's compressed size at 4 K / 8 K / 64 K (chunks kept raw count in full)"""
    ref, blob = golden["ref"], golden["blob"]
    for ent in ref["files"]:
        d = blob[ent["offset"]: ent["offset"] + ent["size"]]
        for chunk in (4096, 8192, 65536):
            rec = ref["corpus"][ent["name"]][str(chunk)]
            if rec["n"] == 0:
                continue
            blocks, lens = O.lzf_fast_chunks(d, chunk)
            full = _raw_where_zero(d, chunk, blocks, lens)
            if O.have_ref_cli():
                back = O.ref_decompress_chunks(O.COMP_LZF, chunk, full, lens)
                assert (back == d[: len(back)]).all() and len(back) == rec["n"] * chunk
            else:
                assert O.verify_roundtrip(O.COMP_LZF, d, chunk, full, lens) == 0
            stored = lambda ls: sum(int(x) if x else chunk for x in ls)
            assert stored(lens) <= 1.01 * stored(rec["lzf_lens"]), (ent["name"], chunk, stored(lens), stored(rec["lzf_lens"]))


def test_lzf_warp_algorithm_byte_identical(emu, golden):
    if not hasattr(emu, "emu_lzf_compress"):
        pytest.skip("LZF warp kernel not built yet")
    for c in _cases(golden, small=True):
        n = len(c)
        out = np.zeros(n + 64, np.uint8)
        blocks, lens = O.compress_chunks(O.COMP_LZF, c, n)
        # every table format the size allows (lzf_warp.cuh: 12, 13, 25 = 4096 buckets + 8192-entry ring, 0 = general)
        for tagged in ([12, 13, 25, 24, 0] if n <= 4096 else [13, 25, 24, 0] if n <= 8192 else [0]):
            out[:] = 0
            r = emu.emu_lzf_compress_mode(c.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), n - 1, tagged)
            assert r == lens[0], (n, tagged, r, lens[0])
            assert (out[:r] == blocks[0, :r]).all(), (n, tagged)


def test_warp_decoders_restore_reference_streams(emu, golden):
    """csrc/decode_warp.cuh on the emulator: decodes what the oracle (== reference) compressors produced; junk never crashes"""
    emu.emu_decode.restype = C.c_uint32
    rng = np.random.default_rng(1)
    blob = golden["blob"]
    cases = [blob[i * 8192:(i + 1) * 8192].copy() for i in range(0, 70, 5)] + [
        np.zeros(8192, np.uint8), rng.integers(0, 256, 8192, dtype=np.uint8), O.synth(8192, 4, 1), blob[:65536].copy(),
        np.tile(rng.integers(0, 256, 3, dtype=np.uint8), 3000)[:8192], blob[:100].copy()]
    for c in cases:
        n = len(c)
        for alg in (O.COMP_LZ4, O.COMP_LZF):
            blocks, lens = O.compress_chunks(alg, c, n)
            if lens[0] == 0:
                continue
            out = np.zeros(n, np.uint8)
            r = emu.emu_decode(alg, blocks[0].ctypes.data_as(C.c_void_p), int(lens[0]), out.ctypes.data_as(C.c_void_p), n)
            assert r == n and (out == c).all(), (alg, n)
    for alg in (0, 1):
        for _ in range(100):
            junk = rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8)
            out = np.zeros(512, np.uint8)
            emu.emu_decode(alg, junk.ctypes.data_as(C.c_void_p), len(junk), out.ctypes.data_as(C.c_void_p), 512)
