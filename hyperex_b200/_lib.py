"""ctypes binding of the C ABI in include/hyperex_b200.h (hyperex_b200/_lib/libhyperex_b200.so).

The shared library is built in-tree by `make -C hyperex_b200/csrc` (see __graft_entry__.build()).
There is no CPU fallback: if the library is missing, loading raises; if no CUDA device is usable,
hx_create fails and Context() raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HYPEREX_B200_LIB") or os.path.join(_HERE, "_lib", "libhyperex_b200.so")

# every symbol include/hyperex_b200.h declares
EXPORTS = [
    "hx_create", "hx_destroy", "hx_last_error", "hx_abi_version", "hx_set_primers", "hx_set_option", "hx_run_batch",
    "hx_run_batch_device", "hx_find_all_end", "hx_alloc_pinned", "hx_free_pinned", "hx_launch_count",
    "hx_kernel_time", "hx_sequence_type", "hx_to_complement", "hx_to_reverse_complement", "hx_primers_to_region",
    "hx_region_to_primer", "hx_supported_regions", "hx_file_to_vec", "hx_free_primer_file",
    "hx_get_hypervar_regions", "hx_fasta_open", "hx_fasta_next", "hx_fasta_close", "hx_int_peak", "hx_group_info", "hx_plan_groups", "hx_plan_peq_word", "hx_plan_packing", "hx_run_batch_fasta", "hx_run_batch_text",
    "hx_run_batch_text_stream", "hx_file_stats", "hx_batch_stats", "hx_fasta_descs", "hx_pack_records", "hx_run_batch_packed",
    "hx_fasta_configure", "hx_fasta_next_packed",
]

LOG_FN = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)
TEXT_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C hyperex_b200/csrc` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, cp, u8p = C.c_void_p, C.c_char_p, C.POINTER(C.c_uint8)
    L.hx_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.hx_destroy.argtypes = [vp]
    L.hx_destroy.restype = None
    L.hx_last_error.argtypes = [vp]
    L.hx_last_error.restype = cp
    L.hx_abi_version.restype = C.c_int
    L.hx_set_primers.argtypes = [vp, C.c_uint32, C.POINTER(cp), C.POINTER(C.c_uint32), C.POINTER(cp),
                                 C.POINTER(C.c_uint32), C.c_uint8]
    L.hx_set_option.argtypes = [vp, cp, C.c_int64]
    L.hx_run_batch.argtypes = [vp, vp, vp, C.c_uint32, vp]
    L.hx_run_batch_device.argtypes = [vp, C.c_int, vp, vp, C.c_uint32, C.c_uint64, vp, vp]
    L.hx_find_all_end.argtypes = [vp, cp, C.c_uint32, vp, C.c_uint64, C.c_uint8, C.c_int, vp, vp, C.c_uint64]
    L.hx_find_all_end.restype = C.c_int64
    L.hx_alloc_pinned.argtypes = [C.c_size_t]
    L.hx_alloc_pinned.restype = vp
    L.hx_free_pinned.argtypes = [vp]
    L.hx_free_pinned.restype = None
    L.hx_launch_count.argtypes = [vp]
    L.hx_launch_count.restype = C.c_uint64
    L.hx_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.hx_sequence_type.argtypes = [cp, C.c_size_t]
    L.hx_to_complement.argtypes = [cp, C.c_size_t, C.c_int, cp]
    L.hx_to_complement.restype = None
    L.hx_to_reverse_complement.argtypes = [cp, C.c_size_t, C.c_int, cp]
    L.hx_to_reverse_complement.restype = None
    L.hx_primers_to_region.argtypes = [cp, cp, cp]
    L.hx_primers_to_region.restype = None
    L.hx_region_to_primer.argtypes = [cp, C.POINTER(cp), C.POINTER(cp)]
    L.hx_supported_regions.argtypes = [C.POINTER(C.c_uint32)]
    L.hx_supported_regions.restype = C.POINTER(cp)
    L.hx_file_to_vec.argtypes = [cp, C.POINTER(C.POINTER(cp)), C.POINTER(C.POINTER(cp))]
    L.hx_free_primer_file.argtypes = [C.POINTER(cp), C.POINTER(cp), C.c_int]
    L.hx_free_primer_file.restype = None
    L.hx_get_hypervar_regions.argtypes = [vp, cp, C.c_uint32, C.POINTER(cp), C.POINTER(cp), cp, C.c_uint8, LOG_FN,
                                          vp]
    L.hx_fasta_open.argtypes = [cp, C.POINTER(vp)]
    L.hx_fasta_next.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.POINTER(cp))]
    L.hx_fasta_next.restype = C.c_int64
    L.hx_fasta_configure.argtypes = [vp, C.c_int, C.c_int]
    L.hx_fasta_next_packed.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                       C.POINTER(C.POINTER(cp)), C.POINTER(vp)]
    L.hx_fasta_next_packed.restype = C.c_int64
    L.hx_fasta_descs.argtypes = [vp]
    L.hx_fasta_descs.restype = C.POINTER(cp)
    L.hx_fasta_close.argtypes = [vp]
    L.hx_fasta_close.restype = None
    L.hx_int_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.hx_group_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.hx_plan_groups.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_uint8, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp]
    L.hx_plan_peq_word.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_uint8, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                   C.c_int, C.c_uint8, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint32)]
    L.hx_plan_packing.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_uint8, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_uint8,
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.hx_run_batch_fasta.argtypes = [vp, vp, vp, C.c_uint32, vp, cp, vp, cp, vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.hx_run_batch_text_stream.argtypes = [vp, vp, vp, C.c_uint32, vp, cp, vp, cp, vp, cp, vp, TEXT_SINK, vp,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.hx_run_batch_text.argtypes = [vp, vp, vp, C.c_uint32, vp, cp, vp, cp, vp, cp, vp, C.POINTER(vp), C.POINTER(C.c_uint64),
                                    C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.hx_file_stats.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    L.hx_batch_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
    L.hx_pack_records.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_int]
    L.hx_run_batch_packed.argtypes = [vp, vp, vp, vp, C.c_uint32, vp]
    _lib = L
    return L
