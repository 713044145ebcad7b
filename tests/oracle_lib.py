"""ctypes access to the CPU oracle (oracle/liboracle.so) and, where present, to the artefacts built from
the reference itself (oracle/_ref/libskref.so, oracle/_ref/ref_cli).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package (compute_war_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
CORPUS_DIR = os.path.join(REF_DIR, "corpus")

HASH_SKEIN512, HASH_SKEIN256_128, HASH_SHA256 = 0, 1, 2
COMP_LZ4, COMP_LZF = 0, 1
HASH_NAMES = {HASH_SKEIN512: "skein512", HASH_SKEIN256_128: "skein256-128", HASH_SHA256: "sha256"}
COMP_NAMES = {COMP_LZ4: "lz4", COMP_LZF: "lzf"}
DIGEST_BYTES = {HASH_SKEIN512: 64, HASH_SKEIN256_128: 16, HASH_SHA256: 32}

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def build_oracle() -> str:
    """compile oracle/liboracle.so if missing or stale (gcc only; a few hundred ms)"""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("cw_oracle.c", "cw_oracle.h")] + [
        os.path.join(ROOT, "include", "cw_synth.h")
    ]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.cwo_skein.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_uint64, C.c_void_p]
        L.cwo_sha256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.cwo_lz4_compress.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.cwo_lz4_compress.restype = C.c_int
        L.cwo_lz4_fast_compress.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.cwo_lz4_fast_compress.restype = C.c_int
        L.cwo_lz4_fast_hashlog.argtypes = [C.c_size_t]
        L.cwo_lz4_fast_hashlog.restype = C.c_int
        L.cwo_lzf_fast_compress.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.cwo_lzf_fast_compress.restype = C.c_int
        L.cwo_lzf_fast_hashlog.argtypes = [C.c_size_t]
        L.cwo_lzf_fast_hashlog.restype = C.c_int
        L.cwo_lz4_decompress.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.cwo_lz4_decompress.restype = C.c_int
        L.cwo_lzf_compress.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
        L.cwo_lzf_compress.restype = C.c_uint
        L.cwo_lzf_decompress.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
        L.cwo_lzf_decompress.restype = C.c_uint
        L.cwo_compress_bound.argtypes = [C.c_int, C.c_size_t]
        L.cwo_compress_bound.restype = C.c_size_t
        L.cwo_hash_chunks.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.cwo_compress_chunks.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                          C.c_void_p]
        L.cwo_hash_compress_chunks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                               C.c_void_p, C.c_size_t, C.c_void_p]
        L.cwo_verify_roundtrip.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p]
        L.cwo_verify_roundtrip.restype = C.c_int
        L.cwo_hash_compress_mt.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.cwo_hash_compress_mt.restype = C.c_double
        L.cwo_synth_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_size_t, C.c_uint64]
        L.cwo_digest_byte_stats.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        L.cwo_digest_byte_stats.restype = C.c_double
        L.cwo_dedup_first.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.cwo_dedup_first.restype = C.c_size_t
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8).copy()


def skein(state_bits: int, hash_bits: int, msg: bytes, msg_bits: int | None = None) -> bytes:
    m = _as_u8(msg) if len(msg) else np.zeros(1, np.uint8)
    bits = len(msg) * 8 if msg_bits is None else msg_bits
    out = np.zeros((hash_bits + 7) // 8, np.uint8)
    lib().cwo_skein(state_bits, hash_bits, _ptr(m), bits, _ptr(out))
    return out.tobytes()


def sha256(msg: bytes) -> bytes:
    m = _as_u8(msg) if len(msg) else np.zeros(1, np.uint8)
    out = np.zeros(32, np.uint8)
    lib().cwo_sha256(_ptr(m), len(msg), _ptr(out))
    return out.tobytes()


def compress_bound(comp: int, chunk: int) -> int:
    return int(lib().cwo_compress_bound(comp, chunk))


def hash_chunks(alg: int, data, chunk: int) -> np.ndarray:
    d = _as_u8(data)
    n = len(d) // chunk
    out = np.zeros((n, DIGEST_BYTES[alg]), np.uint8)
    if n:
        lib().cwo_hash_chunks(alg, _ptr(d), chunk, n, _ptr(out))
    return out


def lz4_fast_chunks(data, chunk: int, stride: int | None = None, hashlog: int = 0):
    """the ENGINE's fast-mode LZ4 parse restated on the CPU (not a reference algorithm): (blocks[n, stride], lens[n])"""
    d = _as_u8(data)
    n = len(d) // chunk
    stride = stride or compress_bound(COMP_LZ4, chunk)
    hashlog = hashlog or int(lib().cwo_lz4_fast_hashlog(chunk))
    blocks = np.zeros((n, stride), np.uint8)
    lens = np.zeros(n, np.uint32)
    for i in range(n):
        lens[i] = lib().cwo_lz4_fast_compress(d[i * chunk:].ctypes.data_as(C.c_void_p), chunk,
                                              blocks[i].ctypes.data_as(C.c_void_p), stride, hashlog)
    return blocks, lens


def lzf_fast_chunks(data, chunk: int, stride: int | None = None, hashlog: int = 0):
    """the ENGINE's fast-mode LZF parse restated on the CPU (not a reference algorithm): (blocks[n, stride], lens[n]);
    lens 0 = the block would not fit in chunk - 1 bytes (the chunk is kept raw)"""
    d = _as_u8(data)
    n = len(d) // chunk
    stride = stride or compress_bound(COMP_LZF, chunk)
    hashlog = hashlog or int(lib().cwo_lzf_fast_hashlog(chunk))
    blocks = np.zeros((n, stride), np.uint8)
    lens = np.zeros(n, np.uint32)
    for i in range(n):
        lens[i] = lib().cwo_lzf_fast_compress(d[i * chunk:].ctypes.data_as(C.c_void_p), chunk,
                                              blocks[i].ctypes.data_as(C.c_void_p), chunk - 1, hashlog)
    return blocks, lens


def compress_chunks(alg: int, data, chunk: int, stride: int | None = None):
    """-> (blocks[n, stride] u8, lens[n] u32) with the reference's per-codec conventions"""
    d = _as_u8(data)
    n = len(d) // chunk
    stride = stride or compress_bound(alg, chunk)
    blocks = np.zeros((n, stride), np.uint8)
    lens = np.zeros(n, np.uint32)
    if n:
        lib().cwo_compress_chunks(alg, _ptr(d), chunk, n, _ptr(blocks), stride, _ptr(lens))
    return blocks, lens


def verify_roundtrip(alg: int, data, chunk: int, blocks: np.ndarray, lens: np.ndarray) -> int:
    """0 if every chunk decodes back to the input, else 1 + index of the first bad chunk"""
    d = _as_u8(data)
    n = len(lens)
    b = np.ascontiguousarray(blocks, dtype=np.uint8)
    l = np.ascontiguousarray(lens, dtype=np.uint32)
    stride = b.shape[1] if b.ndim == 2 else b.size // max(n, 1)
    return int(lib().cwo_verify_roundtrip(alg, _ptr(d), chunk, n, _ptr(b), stride, _ptr(l)))


def synth(chunk: int, first_chunk: int, n: int, seed: int = 0xC0FFEE) -> np.ndarray:
    out = np.zeros(n * chunk, np.uint8)
    lib().cwo_synth_fill(_ptr(out), chunk, first_chunk, n, seed)
    return out


def digest_byte_stats(digests: np.ndarray, n_bytes: int = 16):
    """(avg[n_bytes], stddev[n_bytes], average std dev) as src/hashing_randomness/hash.cpp:30-101 logs them"""
    d = np.ascontiguousarray(digests, dtype=np.uint8)
    n, D = d.shape
    avg = np.zeros(n_bytes, dtype=np.float64)
    sd = np.zeros(n_bytes, dtype=np.float64)
    tot = lib().cwo_digest_byte_stats(_ptr(d), D, n, n_bytes, _ptr(avg), _ptr(sd))
    return avg, sd, float(tot)


def dedup_first(digests: np.ndarray):
    d = np.ascontiguousarray(digests, dtype=np.uint8)
    n, D = d.shape
    first = np.zeros(max(n, 1), dtype=np.uint32)
    uniq = lib().cwo_dedup_first(_ptr(d), D, n, _ptr(first))
    return first[:n], int(uniq)


def hash_compress_mt(halg: int, calg: int, data: np.ndarray, chunk: int, threads: int):
    d = _as_u8(data)
    n = len(d) // chunk
    stride = compress_bound(calg, chunk)
    dig = np.zeros((n, DIGEST_BYTES[halg]), np.uint8)
    blocks = np.zeros((n, stride), np.uint8)
    lens = np.zeros(n, np.uint32)
    sec = lib().cwo_hash_compress_mt(halg, calg, _ptr(d), chunk, n, _ptr(dig), _ptr(blocks), stride, _ptr(lens),
                                     threads)
    return sec, dig, blocks, lens


# ---------------------------------------------------------------------------------------------
# artefacts built from the reference itself (present when `make -C oracle ref` ran in the build container)
# ---------------------------------------------------------------------------------------------

def have_ref_cli() -> bool:
    return os.access(os.path.join(REF_DIR, "ref_cli"), os.X_OK)


def have_skref() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libskref.so"))


def ref_cli(*args: str) -> str:
    return subprocess.check_output([os.path.join(REF_DIR, "ref_cli"), *args], text=True)


_REF_BENCH_EXE = None


def ref_cli_bench_exe() -> str:
    """the driver whose Skein is compiled for x86-64-v3 (oracle/Makefile) if this host runs it, else the portable one"""
    global _REF_BENCH_EXE
    if _REF_BENCH_EXE is None:
        _REF_BENCH_EXE = os.path.join(REF_DIR, "ref_cli")
        v3 = os.path.join(REF_DIR, "ref_cli_v3")
        try:
            if os.access(v3, os.X_OK) and subprocess.run([v3, "bench", "skein512", "lz4", "8192", "64", "1", "0"],
                                                          capture_output=True, timeout=60).returncode == 0:
                _REF_BENCH_EXE = v3
        except Exception:
            pass
    return _REF_BENCH_EXE


def ref_hash_chunks(alg: int, data, chunk: int) -> np.ndarray:
    d = _as_u8(data)
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in"), os.path.join(td, "dig")
        d.tofile(fi)
        ref_cli("hash", HASH_NAMES[alg], str(chunk), fi, fo)
        return np.fromfile(fo, np.uint8).reshape(-1, DIGEST_BYTES[alg])


def ref_compress_chunks(alg: int, data, chunk: int):
    d = _as_u8(data)
    with tempfile.TemporaryDirectory() as td:
        fi, fb, fl = (os.path.join(td, x) for x in ("in", "blocks", "lens"))
        d.tofile(fi)
        ref_cli("compress", COMP_NAMES[alg], str(chunk), fi, fb, fl)
        lens = np.fromfile(fl, np.uint32)
        blocks = np.fromfile(fb, np.uint8).reshape(len(lens), -1) if len(lens) else np.zeros((0, 0), np.uint8)
        return blocks, lens


def ref_decompress_chunks(alg: int, chunk: int, blocks: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """decode with the reference's own LZ4_decompress_safe / lzf_decompress; raises if any chunk fails"""
    with tempfile.TemporaryDirectory() as td:
        fb, fl, fo = (os.path.join(td, x) for x in ("blocks", "lens", "out"))
        np.ascontiguousarray(blocks, np.uint8).tofile(fb)
        np.ascontiguousarray(lens, np.uint32).tofile(fl)
        ref_cli("decompress", COMP_NAMES[alg], str(chunk), fb, fl, fo)
        return np.fromfile(fo, np.uint8)


class SkRef:
    """the reference's Skein source compiled as-is (oracle/_ref/libskref.so)"""

    def __init__(self):
        self.L = C.CDLL(os.path.join(REF_DIR, "libskref.so"))

    def _run(self, name: str, hash_bits: int, msg: bytes) -> bytes:
        ctx = C.create_string_buffer(512)
        m = C.create_string_buffer(bytes(msg), max(len(msg), 1))
        out = C.create_string_buffer((hash_bits + 7) // 8 + 64)
        getattr(self.L, f"{name}_Init")(ctx, C.c_size_t(hash_bits))
        getattr(self.L, f"{name}_Update")(ctx, m, C.c_size_t(len(msg)))
        getattr(self.L, f"{name}_Final")(ctx, out)
        return out.raw[: (hash_bits + 7) // 8]

    def skein512(self, msg: bytes, hash_bits: int = 512) -> bytes:
        return self._run("Skein_512", hash_bits, msg)

    def skein256(self, msg: bytes, hash_bits: int = 128) -> bytes:
        return self._run("Skein_256", hash_bits, msg)


def corpus_files() -> list[str]:
    """full corpus staged by `make -C oracle ref` (git-ignored, travels with gpurun); may be empty"""
    out = []
    for sub in ("canterbury", "canterbury-large"):
        d = os.path.join(CORPUS_DIR, sub)
        if os.path.isdir(d):
            out += [os.path.join(d, f) for f in sorted(os.listdir(d))]
    return out
