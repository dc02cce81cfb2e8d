"""ctypes binding of the C ABI in include/goetia_b200.h (libgoetia_b200.so).

This is the ONLY route from Python to the compute path.  There is no CPU fallback: if the
library is missing, or no sm_100 GPU is visible when a compute call is made, it raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgoetia_b200.so")

u64p = C.POINTER(C.c_uint64)
i16p = C.POINTER(C.c_int16)
u8p = C.POINTER(C.c_uint8)

STORAGE_BIT, STORAGE_BYTE, STORAGE_NIBBLE = 0, 1, 2
SHIFTER_FWD, SHIFTER_CAN = 0, 1
MODE_BLIND, MODE_FAST, MODE_EXACT = 0, 1, 2
READ_OK, READ_SHORT, READ_INVALID = 0, 1, 2

# every symbol include/goetia_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "gt_abi_version": (C.c_int, []),
    "gt_init": (C.c_int, [C.c_int]),
    "gt_device_count": (C.c_int, []),
    "gt_last_error": (C.c_char_p, []),
    "gt_synchronize": (C.c_int, []),
    "gt_primes_near": (C.c_int, [C.c_uint32, C.c_uint64, u64p]),
    "gt_storage_create": (C.c_void_p, [C.c_int, u64p, C.c_int]),
    "gt_storage_destroy": (None, [C.c_void_p]),
    "gt_storage_reset": (C.c_int, [C.c_void_p]),
    "gt_storage_kind": (C.c_int, [C.c_void_p]),
    "gt_storage_n_tables": (C.c_int, [C.c_void_p]),
    "gt_storage_tablesizes": (C.c_int, [C.c_void_p, u64p]),
    "gt_storage_table_bytes": (C.c_uint64, [C.c_void_p, C.c_int]),
    "gt_storage_download_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gt_storage_upload_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gt_storage_stats": (C.c_int, [C.c_void_p, u64p, u64p]),
    "gt_storage_set_n_unique": (C.c_int, [C.c_void_p, C.c_uint64]),
    "gt_storage_checksum": (C.c_int, [C.c_void_p, C.c_int, u64p]),
    "gt_storage_update_from": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_storage_device_table": (C.c_void_p, [C.c_void_p, C.c_int]),
    "gt_insert_hashes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "gt_query_hashes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gt_hash_sequences": (C.c_int64, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "gt_insert_sequences": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "gt_query_sequences": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                       C.c_void_p]),
    "gt_median_count_at_least": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                             C.c_uint32, C.c_void_p, C.c_void_p]),
    "gt_diginorm_sequences": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_uint32, C.c_void_p, C.c_void_p]),
    "gt_diginorm_sequences_serial": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                                 C.c_uint32, C.c_void_p, C.c_void_p]),
    "gt_insert_and_query_sequences": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                  C.c_void_p]),
    "gt_batch_pack": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_batch_pack_dev": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "gt_batch_destroy": (None, [C.c_void_p]),
    "gt_batch_n_reads": (C.c_uint64, [C.c_void_p]),
    "gt_batch_n_bases": (C.c_uint64, [C.c_void_p]),
    "gt_batch_n_kmers": (C.c_int64, [C.c_void_p, C.c_int]),
    "gt_batch_status": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gt_insert_batch": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gt_insert_sequences_dev": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.c_uint64, C.c_int]),
    "gt_insert_sequences_dev_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                                C.c_uint64, C.c_int, C.c_void_p]),
    "gt_pack_reads_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "gt_insert_sequences_packed": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                               C.c_int]),
    "gt_insert_packed_dev_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]),
    "gt_storage_flush": (C.c_int, [C.c_void_p]),
    "gt_storage_pending_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_storage_apply": (C.c_int, [C.c_void_p]),
    "gt_shard_plan": (C.c_int, [C.c_int, u64p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gt_storage_create_sharded": (C.c_void_p, [C.c_int, u64p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "gt_storage_local_range": (C.c_int, [C.c_void_p, C.c_int, u64p, u64p]),
    "gt_storage_attach_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gt_peer_alloc": (C.c_void_p, [C.c_uint64]),
    "gt_peer_free": (C.c_int, [C.c_void_p]),
    "gt_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_peer_open": (C.c_void_p, [C.c_void_p]),
    "gt_peer_close": (C.c_int, [C.c_void_p]),
    "gt_shard_peer_layout": (C.c_int, [C.c_int, u64p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "gt_storage_inbox_bytes": (C.c_uint64, [C.c_void_p, C.c_int]),
    "gt_storage_attach_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gt_storage_stage_bytes": (C.c_uint64, [C.c_void_p, C.c_int]),
    "gt_storage_attach_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gt_storage_attach_areas": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gt_peer_copy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gt_storage_hint_kmers": (C.c_int, [C.c_void_p, C.c_uint64]),
    "gt_query_hashes_local_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gt_shard_route_hashes_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "gt_shard_answer_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gt_shard_insert_requests_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gt_hash_values_dev": (C.c_int64, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "gt_storage_select_store": (C.c_int, [C.c_void_p, C.c_int]),
    "gt_storage_apply_store": (C.c_int, [C.c_void_p, C.c_int]),
    "gt_set_compute_stream": (C.c_int, [C.c_void_p]),
    "gt_set_apply_stream": (C.c_int, [C.c_void_p]),
    "gt_launch_count": (C.c_uint64, []),
    "gt_synth_bases_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "gt_probe_random": (C.c_double, [C.c_int, C.c_uint64, C.c_uint64]),
    "gt_median_count_at_least_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                               C.c_uint32, C.c_void_p, C.c_void_p]),
    "gt_timer_record": (C.c_int, [C.c_int]),
    "gt_timer_elapsed_ms": (C.c_double, [C.c_int, C.c_int]),
    "gt_profile_enable": (C.c_int, [C.c_int]),
    "gt_profile_get": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_profile_get_detail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gt_fastx_open": (C.c_void_p, [C.c_char_p, C.c_int, C.c_uint32]),
    "gt_fastx_close": (None, [C.c_void_p]),
    "gt_fastx_next_record": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "gt_fastx_next_batch": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "gt_fastx_next_packed_batch": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "gt_fastx_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gt_insert_fastx": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]),
    "gt_insert_fastx_advance": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "gt_max_hash_from_scaled": (C.c_uint64, [C.c_uint64]),
    "gt_sketch_create": (C.c_void_p, [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64]),
    "gt_sketch_destroy": (None, [C.c_void_p]),
    "gt_sketch_add_sequences": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_sketch_add_hashes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_sketch_live_count": (C.c_int64, [C.c_void_p]),
    "gt_sketch_export_dev": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_sketch_add_hashes_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_sketch_merge": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gt_sketch_size": (C.c_int64, [C.c_void_p]),
    "gt_sketch_mins": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "gt_sketch_add_sequences_dev": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "gt_sketch_count_common": (C.c_int64, [C.c_void_p, C.c_void_p]),
    "gt_sketch_reset": (C.c_int, [C.c_void_p]),
}


class GoetiaB200Error(RuntimeError):
    """Raised for every failure reported through the C ABI (stands in for GoetiaException)."""


_lib = None
_device = None


def load():
    """dlopen the library and bind every declared symbol (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GoetiaB200Error(
                "%s is missing: build it with `python -m goetia_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    msg = load().gt_last_error()
    return msg.decode() if msg else ""


def check(rc, what=""):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise GoetiaB200Error("%s: %s" % (what, last_error()))
    return rc


def init(device=None):
    """Bind this process to one GPU (LOCAL_RANK by default).  Raises when no GPU is usable."""
    global _device
    L = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _device is None else _device
    if _device is not None and _device == device:
        return device
    n = L.gt_device_count()
    if n <= 0:
        raise GoetiaB200Error("no CUDA device visible: the goetia_b200 compute path needs a B200 (no CPU fallback)")
    check(L.gt_init(int(device)), "gt_init")
    _device = device
    return device


def lib():
    """The bound library with a device selected."""
    L = load()
    if _device is None:
        init()
    return L


def as_reads(bases, offsets):
    """Normalise (bases, offsets) to contiguous uint8 / uint64 numpy arrays."""
    if isinstance(bases, (bytes, bytearray)):
        bases = np.frombuffer(bases, dtype=np.uint8)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if offsets.ndim != 1 or offsets.size < 1:
        raise ValueError("offsets must be a 1-D array of n_reads+1 entries")
    if offsets.size > 1 and int(offsets[-1]) > bases.size:
        raise ValueError("offsets[-1] exceeds len(bases)")
    return bases, offsets


def reads_from_strings(seqs):
    """List of str/bytes -> (bases, offsets)."""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(bases), offsets
