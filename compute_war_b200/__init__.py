"""compute_war_b200 -- B200-native chunked hash-and-compress engine (Skein / SHA-256 x LZ4 / LZF).

Drop-in for the per-chunk path of ankitkumar93/compute-war (src/hashandcompress, src/hashing_perf,
src/compression_perf lz4/lzf branches).  The product is the CUDA library `libcomputewar.so` behind the C ABI of
include/compute_war.h; this package is the thin ctypes binding used by the tests and bench.py.
There is no CPU fallback: importing works anywhere, creating an Engine needs a CUDA device.
"""
from .engine import (COMP_LZ4, COMP_LZF, HASH_SHA256, HASH_SKEIN256_128, HASH_SKEIN512, MODE_COMPAT, MODE_FAST,
                     Engine, EngineError, compress_bound, device_count, digest_bytes, library_path, load_library)

__all__ = ["Engine", "EngineError", "load_library", "library_path", "digest_bytes", "compress_bound", "device_count",
           "HASH_SKEIN512", "HASH_SKEIN256_128", "HASH_SHA256", "COMP_LZ4", "COMP_LZF", "MODE_COMPAT", "MODE_FAST"]
