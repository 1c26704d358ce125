"""ctypes wrapper around the CPU oracle (oracle/hx_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (hyperex_b200/) never imports it.
PARITY STATUS: "parity unpinned" by the reference's tests; pinned by rust-bio
documentation vectors and an independent DP (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhxoracle.so")


class HxoResult(C.Structure):
    _fields_ = [("f_start", C.c_uint32), ("r_end", C.c_uint32), ("flags", C.c_uint8),
                ("combined", C.c_uint8), ("reserved", C.c_uint16)]


RESULT_DTYPE = np.dtype([("f_start", "<u4"), ("r_end", "<u4"), ("flags", "u1"), ("combined", "u1"),
                         ("reserved", "<u2")])


class HxoMyers(C.Structure):
    _fields_ = [("peq", C.c_uint64 * 256), ("bound", C.c_uint64), ("m", C.c_uint32)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "hx_oracle.c"), os.path.join(_HERE, "hx_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        L.hxo_myers_build.argtypes = [C.POINTER(HxoMyers), C.c_char_p, C.c_uint32, C.c_int]
        L.hxo_myers_build_custom.argtypes = [C.POINTER(HxoMyers), C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p,
                                             C.POINTER(C.c_char_p)]
        L.hxo_find_all_end.restype = C.c_size_t
        L.hxo_find_all_end.argtypes = [C.POINTER(HxoMyers), C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p, C.c_void_p,
                                       C.c_size_t]
        L.hxo_dp_find_all_end.restype = C.c_size_t
        L.hxo_dp_find_all_end.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p,
                                          C.c_void_p, C.c_size_t]
        L.hxo_sequence_type.argtypes = [C.c_char_p, C.c_size_t]
        L.hxo_to_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p]
        L.hxo_to_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p]
        L.hxo_primers_to_region.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.hxo_region_to_primer.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        L.hxo_supported_regions.restype = C.POINTER(C.c_char_p)
        L.hxo_supported_regions.argtypes = [C.POINTER(C.c_uint32)]
        L.hxo_pair_literal.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_size_t, u64p, u64p, u8p]
        L.hxo_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p),
                                    C.POINTER(C.c_char_p), C.c_uint8, C.c_void_p, C.c_int, C.c_int]
        L.hxo_format_outputs.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_void_p,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_size_t)]
        L.hxo_free.argtypes = [C.c_void_p]
        L.hxo_parse_fasta.restype = C.c_int64
        L.hxo_parse_fasta.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


DNA, RNA, NONE = 0, 1, 2


def _b(s) -> bytes:
    return s.encode() if isinstance(s, str) else bytes(s)


def find_all_end(pattern, text, k: int, ambigs: dict | None = None, hyperex_ambigs: bool = True):
    """[(end, dist)] as Myers<u64>::find_all_lazy(text, k).collect() (utils.rs:305-309)."""
    L = lib()
    pattern, text = _b(pattern), _b(text)
    my = HxoMyers()
    if ambigs is not None:
        codes = bytes(_b(c)[0] for c in ambigs)
        arr = (C.c_char_p * max(1, len(ambigs)))(*[_b(v) for v in ambigs.values()])
        rc = L.hxo_myers_build_custom(C.byref(my), pattern, len(pattern), len(ambigs), codes, arr)
    else:
        rc = L.hxo_myers_build(C.byref(my), pattern, len(pattern), 1 if hyperex_ambigs else 0)
    if rc != 0:
        raise ValueError("pattern length must be in 1..64")
    n = len(text)
    ends = np.empty(max(n, 1), dtype=np.uint64)
    dists = np.empty(max(n, 1), dtype=np.uint8)
    tb = np.frombuffer(text, dtype=np.uint8) if n else np.zeros(1, np.uint8)
    cnt = L.hxo_find_all_end(C.byref(my), tb.ctypes.data, n, k, ends.ctypes.data, dists.ctypes.data, n)
    return [(int(ends[i]), int(dists[i])) for i in range(cnt)]


def dp_find_all_end(pattern, text, k: int):
    L = lib()
    pattern, text = _b(pattern), _b(text)
    n = len(text)
    ends = np.empty(max(n, 1), dtype=np.uint64)
    dists = np.empty(max(n, 1), dtype=np.uint8)
    tb = np.frombuffer(text, dtype=np.uint8) if n else np.zeros(1, np.uint8)
    cnt = L.hxo_dp_find_all_end(pattern, len(pattern), tb.ctypes.data, n, k, ends.ctypes.data, dists.ctypes.data, n)
    return [(int(ends[i]), int(dists[i])) for i in range(cnt)]


def sequence_type(seq) -> int:
    s = _b(seq)
    return lib().hxo_sequence_type(s, len(s))


def to_complement(primer, alphabet: int) -> bytes:
    p = _b(primer)
    out = C.create_string_buffer(len(p) + 1)
    lib().hxo_to_complement(p, len(p), alphabet, out)
    return out.raw[:len(p)]


def to_reverse_complement(primer, alphabet: int) -> bytes:
    p = _b(primer)
    out = C.create_string_buffer(len(p) + 1)
    lib().hxo_to_reverse_complement(p, len(p), alphabet, out)
    return out.raw[:len(p)]


def primers_to_region(fwd, rev) -> str:
    out = C.create_string_buffer(8)
    lib().hxo_primers_to_region(_b(fwd), _b(rev), out)
    return out.value.decode()


def region_to_primer(region):
    f, r = C.c_char_p(), C.c_char_p()
    if lib().hxo_region_to_primer(_b(region), C.byref(f), C.byref(r)) != 0:
        return []
    return [f.value.decode(), r.value.decode()]


def supported_regions():
    n = C.c_uint32()
    p = lib().hxo_supported_regions(C.byref(n))
    return [p[i].decode() for i in range(n.value)]


def default_pairs():
    return [region_to_primer(r) for r in supported_regions()]


def pair_literal(f_hits, r_hits, forward_len: int):
    """utils.rs:327-351 on [(end, dist)] lists -> (f_start, r_end, combined) or None."""
    fe = np.array([h[0] for h in f_hits], dtype=np.uint64)
    fd = np.array([h[1] for h in f_hits], dtype=np.uint8)
    re_ = np.array([h[0] for h in r_hits], dtype=np.uint64)
    rd = np.array([h[1] for h in r_hits], dtype=np.uint8)
    fs, rr, cb = C.c_uint64(), C.c_uint64(), C.c_uint8()
    ok = lib().hxo_pair_literal(fe.ctypes.data, fd.ctypes.data, len(fe), re_.ctypes.data, rd.ctypes.data, len(re_),
                                forward_len, C.byref(fs), C.byref(rr), C.byref(cb))
    return (fs.value, rr.value, cb.value) if ok else None


def _cstr_array(strs):
    return (C.c_char_p * max(1, len(strs)))(*[_b(s) for s in strs])


def run_batch(arena: np.ndarray, offsets: np.ndarray, pairs, k: int, nthreads: int = 1,
              rebuild_per_record: bool = False) -> np.ndarray:
    """utils.rs:270-351 over an arena; returns a (n_records, n_pairs) structured array."""
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    out = np.zeros((n, len(pairs)), dtype=RESULT_DTYPE)
    fw = _cstr_array([p[0] for p in pairs])
    rv = _cstr_array([p[1] for p in pairs])
    ap = arena.ctypes.data if arena.size else None
    rc = lib().hxo_run_batch(ap, offsets.ctypes.data, n, len(pairs), fw, rv, k, out.ctypes.data, nthreads,
                             1 if rebuild_per_record else 0)
    if rc != 0:
        raise ValueError("primer length must be in 1..64")
    return out


def format_outputs(arena, offsets, ids, pairs, res):
    """(fasta_bytes, gff_bytes) as written by utils.rs:353-385."""
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    res = np.ascontiguousarray(res)
    n = len(offsets) - 1
    idarr = _cstr_array(ids)
    fw = _cstr_array([p[0] for p in pairs])
    rv = _cstr_array([p[1] for p in pairs])
    fa, gff = C.c_void_p(), C.c_void_p()
    fl, gl = C.c_size_t(), C.c_size_t()
    L = lib()
    L.hxo_format_outputs(arena.ctypes.data if arena.size else None, offsets.ctypes.data, idarr, n, len(pairs), fw,
                         rv, res.ctypes.data, C.byref(fa), C.byref(fl), C.byref(gff), C.byref(gl))
    fab = C.string_at(fa.value, fl.value)
    gfb = C.string_at(gff.value, gl.value)
    L.hxo_free(fa)
    L.hxo_free(gff)
    return fab, gfb


def parse_fasta(buf: bytes):
    """rust-bio fasta::Reader rules -> (arena u8[], offsets u64[n+1], ids [str])."""
    L = lib()
    a, o, i, io = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = L.hxo_parse_fasta(buf, len(buf), C.byref(a), C.byref(o), C.byref(i), C.byref(io))
    offsets = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
    total = int(offsets[-1])
    arena = (np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_uint8)), shape=(max(total, 1),)).copy()[:total])
    ids = []
    if n:
        id_off = np.ctypeslib.as_array(C.cast(io, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        for j in range(n):
            ids.append(C.string_at(i.value + int(id_off[j])).decode("utf-8", "replace"))
    for p in (a, o, i, io):
        L.hxo_free(p)
    return arena, offsets, ids


def get_hypervar_regions(fasta_bytes: bytes, pairs, k: int):
    """Whole reference path on an uncompressed FASTA buffer -> (fa_bytes, gff_bytes, results)."""
    arena, offsets, ids = parse_fasta(fasta_bytes)
    res = run_batch(arena, offsets, pairs, k)
    fa, gff = format_outputs(arena, offsets, ids, pairs, res)
    return fa, gff, res
