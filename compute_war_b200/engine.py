"""ctypes binding of include/compute_war.h.

The names follow the reference's own vocabulary: `do_hashing` / `do_compression` are the two function slots of
src/hashandcompress/HashAndCompress.cpp:111,119 and `process_block` is ProcessBlock (:231-261), here applied to a
whole batch of chunks per call.  Host-buffer methods take numpy arrays / bytes; `*_dev` methods take raw device
pointers (ints, e.g. torch.Tensor.data_ptr()) and a CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HASH_SKEIN512, HASH_SKEIN256_128, HASH_SHA256 = 0, 1, 2
COMP_LZ4, COMP_LZF = 0, 1
MODE_COMPAT, MODE_FAST = 0, 1
MAX_CHUNK = 65536

HASH_NAMES = {"skein512": HASH_SKEIN512, "skein": HASH_SKEIN256_128, "skein256-128": HASH_SKEIN256_128,
              "sha256": HASH_SHA256, "sha256mb": HASH_SHA256}
COMP_NAMES = {"lz4": COMP_LZ4, "lzf": COMP_LZF}

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class EngineError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"compute_war: {msg} (status {status})")
        self.status = status


def library_path() -> str:
    # CW_LIB: alternative build of the same library (A/B measurements of kernel variants); never a CPU fallback
    return os.environ.get("CW_LIB") or os.path.join(_HERE, "libcomputewar.so")


def load_library():
    """Load the in-tree CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(-2, f"{path} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a); there is no CPU fallback")
    L = C.CDLL(path)
    vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
    L.cw_init.argtypes = [C.POINTER(vp), C.c_int]
    L.cw_shutdown.argtypes = [vp]
    L.cw_shutdown.restype = None
    L.cw_last_error.argtypes = [vp]
    L.cw_last_error.restype = C.c_char_p
    L.cw_status_string.argtypes = [C.c_int]
    L.cw_status_string.restype = C.c_char_p
    L.cw_device_count.restype = C.c_int
    L.cw_digest_bytes.argtypes = [C.c_int]
    L.cw_digest_bytes.restype = sz
    L.cw_compress_bound.argtypes = [C.c_int, sz]
    L.cw_compress_bound.restype = sz
    L.cw_hash.argtypes = [vp, C.c_int, vp, sz, sz, vp]
    L.cw_hash_skein.argtypes = [vp, C.c_uint, C.c_uint, vp, sz, sz, vp]
    L.cw_compress.argtypes = [vp, C.c_int, C.c_int, vp, sz, sz, vp, sz, vp]
    L.cw_hash_compress.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, sz, sz, vp, vp, sz, vp]
    L.cw_hash_compress_packed.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, sz, sz, vp, vp, sz, vp, vp]
    L.cw_hash_dev.argtypes = [vp, C.c_int, vp, sz, sz, vp, vp]
    L.cw_compress_dev.argtypes = [vp, C.c_int, C.c_int, vp, sz, sz, vp, sz, vp, vp]
    L.cw_hash_compress_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, sz, sz, vp, vp, sz, vp, vp]
    L.cw_decompress.argtypes = [vp, C.c_int, vp, sz, vp, sz, sz, vp, vp]
    L.cw_decompress_dev.argtypes = [vp, C.c_int, vp, sz, vp, sz, sz, vp, vp, vp]
    L.cw_pack_dev.argtypes = [vp, vp, sz, vp, vp, sz, sz, vp, sz, vp, vp]
    L.cw_digest_hist_dev.argtypes = [vp, vp, sz, sz, vp, vp]
    L.cw_digest_byte_stats.argtypes = [vp, sz, vp, vp, C.POINTER(C.c_double)]
    L.cw_digest_stats.argtypes = [vp, vp, sz, sz, sz, vp, vp, C.POINTER(C.c_double)]
    L.cw_dedup_dev.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.cw_host_register.argtypes = [vp, vp, sz]
    L.cw_host_unregister.argtypes = [vp, vp]
    L.cw_sync.argtypes = [vp, vp]
    L.cw_submit.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, sz, sz, vp, vp, sz, vp, vp, vp, C.POINTER(u64)]
    L.cw_wait.argtypes = [vp, u64]
    L.cw_synth_dev.argtypes = [vp, vp, sz, u64, sz, u64, vp]
    L.cw_batch_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.cw_batch_stats.restype = None
    L.cw_launch_count.argtypes = [vp]
    L.cw_launch_count.restype = u64
    L.cw_prof_enable.argtypes = [vp, C.c_int]
    L.cw_prof_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
    _lib = L
    return L


def device_count() -> int:
    """CUDA devices the library sees (0 without a GPU; raises if the library is not built)"""
    return int(load_library().cw_device_count())


def digest_bytes(hash_alg: int) -> int:
    return {HASH_SKEIN512: 64, HASH_SKEIN256_128: 16, HASH_SHA256: 32}[hash_alg]


def compress_bound(comp_alg: int, chunk: int) -> int:
    """LZ4_COMPRESSBOUND (lz4.h:156) / chunk for LZF (called with out_len = chunk-1)"""
    return chunk + chunk // 255 + 16 if comp_alg == COMP_LZ4 else chunk


def _host_u8(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x)
        return a.view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(x), dtype=np.uint8)


class Engine:
    """One cw_ctx on one GPU (replaces initializeGpu(), HashAndCompress.cpp:95-98)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.cw_init(C.byref(h), device)
        if rc != 0:
            raise EngineError(rc, self.L.cw_status_string(rc).decode() + " " + self.L.cw_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.cw_shutdown(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            detail = self.L.cw_status_string(rc).decode()
            if rc == -3:
                detail += ": " + self.L.cw_last_error(self.h).decode()
            raise EngineError(rc, detail)

    # ---- host-buffer calls (the reference's function slots) ----
    def do_hashing(self, hash_alg: int, src, chunk: int) -> np.ndarray:
        """doHashing(src, dst, count): digests[n, D] of the n = len(src)//chunk whole chunks"""
        s = _host_u8(src)
        n = s.size // chunk
        out = np.zeros((n, digest_bytes(hash_alg)), np.uint8)
        self._check(self.L.cw_hash(self.h, hash_alg, s.ctypes.data, chunk, n, out.ctypes.data))
        return out

    def hash_skein(self, state_bits: int, hash_bits: int, src, chunk: int) -> np.ndarray:
        """Skein-<state_bits> with a digest of hash_bits (Skein_512_Init / Skein_256_Init with any hashBitLen)"""
        s = _host_u8(src)
        n = s.size // chunk
        out = np.zeros((n, hash_bits // 8), np.uint8)
        self._check(self.L.cw_hash_skein(self.h, state_bits, hash_bits, s.ctypes.data, chunk, n, out.ctypes.data))
        return out

    def do_compression(self, comp_alg: int, src, chunk: int, mode: int = MODE_COMPAT, stride: int | None = None):
        """doCompression per chunk: (blocks[n, stride], lens[n]) with the codec's own return convention"""
        s = _host_u8(src)
        n = s.size // chunk
        stride = stride or compress_bound(comp_alg, chunk)
        blocks = np.zeros((n, stride), np.uint8)
        lens = np.zeros(n, np.uint32)
        self._check(self.L.cw_compress(self.h, comp_alg, mode, s.ctypes.data, chunk, n, blocks.ctypes.data, stride,
                                       lens.ctypes.data))
        return blocks, lens

    def process_block(self, hash_alg: int, comp_alg: int, src, chunk: int, mode: int = MODE_COMPAT,
                      out=None):
        """ProcessBlock over a batch: (digests, blocks, lens); `out` may supply preallocated (pinned) arrays"""
        s = _host_u8(src)
        n = s.size // chunk
        stride = compress_bound(comp_alg, chunk)
        if out is None:
            dig = np.zeros((n, digest_bytes(hash_alg)), np.uint8)
            blocks = np.zeros((n, stride), np.uint8)
            lens = np.zeros(n, np.uint32)
        else:
            dig, blocks, lens = out
        self._check(self.L.cw_hash_compress(self.h, hash_alg, comp_alg, mode, s.ctypes.data, chunk, n, dig.ctypes.data,
                                            blocks.ctypes.data, stride, lens.ctypes.data))
        return dig, blocks, lens

    def process_block_packed(self, hash_alg: int, comp_alg: int, src, chunk: int, mode: int = MODE_COMPAT, cap: int | None = None):
        """ProcessBlock with packed output: (digests[n, D], packed bytes, offsets[n + 1] u64, lens[n])"""
        s = _host_u8(src)
        n = s.size // chunk
        cap = n * compress_bound(comp_alg, chunk) if cap is None else cap
        dig = np.zeros((n, digest_bytes(hash_alg)), np.uint8)
        packed = np.zeros(max(cap, 1), np.uint8)
        offsets = np.zeros(n + 1, np.uint64)
        lens = np.zeros(n, np.uint32)
        self._check(self.L.cw_hash_compress_packed(self.h, hash_alg, comp_alg, mode, s.ctypes.data, chunk, n, dig.ctypes.data,
                                                   packed.ctypes.data, cap, offsets.ctypes.data, lens.ctypes.data))
        return dig, packed[: int(offsets[n])], offsets, lens

    def process_block_packed_ptr(self, hash_alg, comp_alg, mode, src_ptr, chunk, n, dig_ptr, packed_ptr, cap, offsets_ptr, lens_ptr):
        self._check(self.L.cw_hash_compress_packed(self.h, hash_alg, comp_alg, mode, src_ptr, chunk, n, dig_ptr, packed_ptr, cap,
                                                   offsets_ptr, lens_ptr))

    def process_block_ptr(self, hash_alg, comp_alg, mode, src_ptr, chunk, n, dig_ptr, blocks_ptr, stride, lens_ptr):
        """ProcessBlock on raw HOST pointers (e.g. pinned torch tensors)"""
        self._check(self.L.cw_hash_compress(self.h, hash_alg, comp_alg, mode, src_ptr, chunk, n, dig_ptr, blocks_ptr,
                                            stride, lens_ptr))

    # ---- asynchronous offload (HashOffload Start / Complete) ----
    def submit(self, hash_alg, comp_alg, mode, src_ptr, chunk, n, dig_ptr, blocks_ptr, stride, lens_ptr) -> int:
        t = C.c_uint64()
        self._check(self.L.cw_submit(self.h, hash_alg, comp_alg, mode, src_ptr, chunk, n, dig_ptr, blocks_ptr, stride,
                                     lens_ptr, None, None, C.byref(t)))
        return t.value

    def wait(self, ticket: int):
        self._check(self.L.cw_wait(self.h, ticket))

    # ---- device-buffer calls ----
    def hash_dev(self, hash_alg, d_src, chunk, n, d_digests, stream=0):
        self._check(self.L.cw_hash_dev(self.h, hash_alg, d_src, chunk, n, d_digests, stream or None))

    def compress_dev(self, comp_alg, mode, d_src, chunk, n, d_dst, stride, d_lens, stream=0):
        self._check(self.L.cw_compress_dev(self.h, comp_alg, mode, d_src, chunk, n, d_dst, stride, d_lens, stream or None))

    def hash_compress_dev(self, hash_alg, comp_alg, mode, d_src, chunk, n, d_digests, d_dst, stride, d_lens, stream=0):
        self._check(self.L.cw_hash_compress_dev(self.h, hash_alg, comp_alg, mode, d_src, chunk, n, d_digests, d_dst, stride,
                                                d_lens, stream or None))

    def decompress_dev(self, comp_alg, d_blocks, stride, d_lens, chunk, n, d_out, d_out_lens, stream=0):
        self._check(self.L.cw_decompress_dev(self.h, comp_alg, d_blocks, stride, d_lens, chunk, n, d_out, d_out_lens,
                                             stream or None))

    def do_decompression(self, comp_alg: int, blocks: np.ndarray, lens: np.ndarray, chunk: int):
        """LZ4_decompress_safe / lzf_decompress per block: (out[n*chunk], out_lens[n]); out_lens[i] == 0 marks a bad block"""
        b = np.ascontiguousarray(blocks, dtype=np.uint8)
        l = np.ascontiguousarray(lens, dtype=np.uint32)
        n = len(l)
        stride = b.shape[1] if b.ndim == 2 else (b.size // max(n, 1))
        out = np.zeros(n * chunk, np.uint8)
        out_lens = np.zeros(n, np.uint32)
        self._check(self.L.cw_decompress(self.h, comp_alg, b.ctypes.data, stride, l.ctypes.data, chunk, n, out.ctypes.data,
                                         out_lens.ctypes.data))
        return out, out_lens

    def pack_dev(self, d_blocks, stride, d_lens, d_src, chunk, n, d_packed, packed_cap, d_offsets, stream=0):
        """strided blocks -> contiguous stream + u64 offsets[n + 1] (blocks of length 0 are stored raw from d_src)"""
        self._check(self.L.cw_pack_dev(self.h, d_blocks, stride, d_lens, d_src or None, chunk, n, d_packed, packed_cap,
                                       d_offsets, stream or None))

    # ---- digest post-processing (src/hashing_randomness/hash.cpp:30-101; dedup map) ----
    def digest_hist_dev(self, d_digests, digest_bytes, n, d_hist, stream=0):
        """d_hist: u64[digest_bytes * 256]; hist[b*256 + v] = digests whose byte b equals v"""
        self._check(self.L.cw_digest_hist_dev(self.h, d_digests, digest_bytes, n, d_hist, stream or None))

    def digest_stats(self, digests: np.ndarray, n_bytes: int = 16):
        """Hasher::LogResults figures: (avg[n_bytes], stddev[n_bytes], average std dev)"""
        d = np.ascontiguousarray(digests, dtype=np.uint8)
        n, D = d.shape
        avg = np.zeros(n_bytes, np.float64)
        sd = np.zeros(n_bytes, np.float64)
        tot = C.c_double()
        self._check(self.L.cw_digest_stats(self.h, d.ctypes.data, D, n, n_bytes, avg.ctypes.data, sd.ctypes.data,
                                           C.byref(tot)))
        return avg, sd, tot.value

    def dedup_dev(self, d_digests, digest_bytes, n, d_first, d_n_unique, stream=0):
        """d_first: u32[n], first[i] = smallest j with an equal digest; d_n_unique: u64"""
        self._check(self.L.cw_dedup_dev(self.h, d_digests, digest_bytes, n, d_first, d_n_unique, stream or None))

    def host_register(self, arr: np.ndarray):
        """page-lock a numpy buffer the caller keeps reusing (cw_host_register)"""
        self._check(self.L.cw_host_register(self.h, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._check(self.L.cw_host_unregister(self.h, arr.ctypes.data))

    def synth_dev(self, d_dst, chunk, first_chunk, n, seed=0xC0FFEE, stream=0):
        self._check(self.L.cw_synth_dev(self.h, d_dst, chunk, first_chunk, n, seed, stream or None))

    def sync(self, stream=0):
        self._check(self.L.cw_sync(self.h, stream or None))

    # ---- introspection ----
    def batch_stats(self):
        """(launches made for small host-buffer calls, calls they served)"""
        b, c = C.c_uint64(0), C.c_uint64(0)
        self.L.cw_batch_stats(self.h, C.byref(b), C.byref(c))
        return int(b.value), int(c.value)

    def launch_count(self) -> int:
        return int(self.L.cw_launch_count(self.h))

    def prof_enable(self, on: bool = True):
        self._check(self.L.cw_prof_enable(self.h, 1 if on else 0))

    def prof_read(self):
        ms, n = C.c_double(), C.c_uint64()
        self._check(self.L.cw_prof_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
