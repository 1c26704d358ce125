"""Host-side mirror of the reference's interface for the primer-search-and-extract path.

Names, argument meaning and error behaviour follow Ebedthan/hyperex v0.2.0 `src/utils.rs`
(region_to_primer 112-131, file_to_vec 134-145, primers_to_region 157-166, to_complement 169-194,
to_reverse_complement 197-199, sequence_type 207-228, get_hypervar_regions 231-414) so the parity
tests read like the reference's own unit tests (src/utils.rs:585-749).  Everything is a thin
ctypes call into the C ABI; the compute runs on the GPU or fails.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib

RESULT_DTYPE = np.dtype([("f_start", "<u4"), ("r_end", "<u4"), ("flags", "u1"), ("combined", "u1"),
                         ("reserved", "<u2")])
FWD_ANY, REV_ANY, PAIR_VALID, ALPHABET_NONE, RNA = 1, 2, 4, 8, 16

SUPPORTED_REGIONS = ["v1v2", "v1v3", "v1v9", "v3v4", "v3v5", "v4", "v4v5", "v5v7", "v6v9", "v7v9"]


class HyperexError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hx error {code}: {msg}")
        self.code = code


class Alphabet:
    Dna = 0
    Rna = 1


def _b(s) -> bytes:
    return s.encode() if isinstance(s, str) else bytes(s)


def _cstrs(strs: Sequence):
    return (C.c_char_p * max(1, len(strs)))(*[_b(s) for s in strs])


# ---- pure host helpers (utils.rs string functions) ---------------------------------------
def sequence_type(sequence):
    """utils.rs:207-228 -> Alphabet.Dna / Alphabet.Rna / None"""
    s = _b(sequence)
    r = _lib.load().hx_sequence_type(s, len(s))
    return None if r == 2 else r


def to_complement(primer, alphabet: str) -> str:
    p = _b(primer)
    out = C.create_string_buffer(len(p) + 1)
    _lib.load().hx_to_complement(p, len(p), {"dna": 0, "rna": 1}.get(alphabet, 2), out)
    return out.raw[:len(p)].decode()


def to_reverse_complement(primer, alphabet: str) -> str:
    p = _b(primer)
    out = C.create_string_buffer(len(p) + 1)
    _lib.load().hx_to_reverse_complement(p, len(p), {"dna": 0, "rna": 1}.get(alphabet, 2), out)
    return out.raw[:len(p)].decode()


def primers_to_region(primers: Sequence[str]) -> str:
    out = C.create_string_buffer(16)
    _lib.load().hx_primers_to_region(_b(primers[0]), _b(primers[1]), out)
    return out.value.decode()


def region_to_primer(region: str):
    """utils.rs:112-131: [forward, reverse] or [] for an unsupported name"""
    f, r = C.c_char_p(), C.c_char_p()
    if _lib.load().hx_region_to_primer(_b(region), C.byref(f), C.byref(r)) != 0:
        return []
    return [f.value.decode(), r.value.decode()]


def supported_regions():
    n = C.c_uint32()
    p = _lib.load().hx_supported_regions(C.byref(n))
    return [p[i].decode() for i in range(n.value)]


def default_primers():
    """utils.rs:513-518: all built-in regions in SUPPORTED_REGIONS order"""
    return [region_to_primer(r) for r in supported_regions()]


def file_to_vec(filename: str):
    """utils.rs:134-145; raises ValueError('Primer file must be comma-separated')"""
    L = _lib.load()
    f, r = C.POINTER(C.c_char_p)(), C.POINTER(C.c_char_p)()
    n = L.hx_file_to_vec(_b(filename), C.byref(f), C.byref(r))
    if n == -5:
        raise FileNotFoundError(filename)
    if n < 0:
        raise ValueError("Primer file must be comma-separated")
    out = [[f[i].decode(), r[i].decode()] for i in range(n)]
    L.hx_free_primer_file(f, r, n)
    return out


PACK_CODES = b"\0ACGTRYSWKMBDHVN"   # code -> letter; 'U' shares code 4 with 'T' (hx_pack_records)
REC_HAS_U, REC_HAS_T, REC_NON_IUPAC = 1, 2, 4


def pack_records(arena: np.ndarray, offsets: np.ndarray, n_threads: int = 0, packed: np.ndarray | None = None,
                 rec_flags: np.ndarray | None = None):
    """hx_pack_records: ASCII arena -> (packed u8[offsets[-1] // 2 + 1], rec_flags u8[n]); nibble i of packed = code of
    base i (low nibble = even index).  Pure host code."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if packed is None:
        packed = np.zeros(int(offsets[-1]) // 2 + 1, dtype=np.uint8)
    if rec_flags is None:
        rec_flags = np.zeros(max(n, 1), dtype=np.uint8)
    ap = arena.ctypes.data if arena is not None and arena.size else None
    rc = _lib.load().hx_pack_records(ap, offsets.ctypes.data, n, packed.ctypes.data, rec_flags.ctypes.data, n_threads)
    if rc != 0:
        raise HyperexError(rc, "hx_pack_records: bad arguments")
    return packed, rec_flags[:n]


def _fasta_open(path: str, threads: int, pack_ahead: bool):
    L = _lib.load()
    h = C.c_void_p()
    rc = L.hx_fasta_open(_b(path), C.byref(h))
    if rc != 0:
        raise HyperexError(rc, f"cannot open {path}")
    if threads != 1 or pack_ahead:
        rc = L.hx_fasta_configure(h, threads, 1 if pack_ahead else 0)
        if rc != 0:
            L.hx_fasta_close(h)
            raise HyperexError(rc, "hx_fasta_configure")
    return L, h


def _view(ptr, n, dtype, copy):
    ct = {np.uint8: C.c_uint8, np.uint64: C.c_uint64}[dtype]
    v = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(n, 1),))[:n]
    return v.copy() if copy else v


def read_fasta(path: str, max_bytes: int = 0, threads: int = 1, copy: bool = True, want_ids: bool = True):
    """fasta::Reader::new(read_file(path)).records() (utils.rs:237-238) -> iterator of
    (arena u8[], offsets u64[n+1], ids [str]) batches.  threads > 1 (0: all but one): chunks are parsed ahead of the
    caller by background threads (hx_fasta_configure), same batches.  copy=False yields views into the reader's pinned
    arenas, valid until the next batch is asked for."""
    L, h = _fasta_open(path, threads, False)
    try:
        while True:
            a, o, ids = C.c_void_p(), C.c_void_p(), C.POINTER(C.c_char_p)()
            n = L.hx_fasta_next(h, max_bytes, C.byref(a), C.byref(o), C.byref(ids))
            if n < 0:
                raise HyperexError(int(n), "read error")
            if n == 0:
                return
            offsets = _view(o, n + 1, np.uint64, copy)
            arena = _view(a, int(offsets[-1]), np.uint8, copy)
            yield arena, offsets, ([ids[i].decode("utf-8", "replace") for i in range(n)] if want_ids else None)
    finally:
        L.hx_fasta_close(h)


def read_fasta_packed(path: str, max_bytes: int = 0, threads: int = 1, pack_ahead: bool = True, copy: bool = True,
                      want_ids: bool = True):
    """hx_fasta_next_packed: iterator of (packed u8[], rec_flags u8[n], offsets u64[n+1], ids, arena u8[]) — the batch as
    the packed arena hx_run_batch_packed takes, plus its ASCII arena.  ids is None when want_ids is false."""
    L, h = _fasta_open(path, threads, pack_ahead)
    try:
        while True:
            pk, fl, o, a, ids = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.POINTER(C.c_char_p)()
            n = L.hx_fasta_next_packed(h, max_bytes, C.byref(pk), C.byref(fl), C.byref(o), C.byref(ids), C.byref(a))
            if n < 0:
                raise HyperexError(int(n), "read error")
            if n == 0:
                return
            offsets = _view(o, n + 1, np.uint64, copy)
            total = int(offsets[-1])
            yield (_view(pk, total // 2 + 1, np.uint8, copy), _view(fl, n, np.uint8, copy), offsets,
                   [ids[i].decode("utf-8", "replace") for i in range(n)] if want_ids else None,
                   _view(a, total, np.uint8, copy))
    finally:
        L.hx_fasta_close(h)


# ---- device path -------------------------------------------------------------------------
class Context:
    """Owns an hx_ctx (streams, device buffers, Peq tables) on one or more GPUs."""

    def __init__(self, devices: Iterable[int] = (0,)):
        self._L = _lib.load()
        devs = list(devices)
        arr = (C.c_int * len(devs))(*devs)
        self._h = C.c_void_p()
        rc = self._L.hx_create(C.byref(self._h), arr, len(devs))
        if rc != 0:
            raise HyperexError(rc, self._L.hx_last_error(None).decode())
        self.n_pairs = 0
        self.pairs = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.hx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise HyperexError(rc, self._L.hx_last_error(self._h).decode())

    def set_option(self, name: str, value: int):
        self._check(self._L.hx_set_option(self._h, _b(name), int(value)))

    def set_primers(self, primers: Sequence[Sequence[str]], mismatch: int):
        fw = [_b(p[0]) for p in primers]
        rv = [_b(p[1]) for p in primers]
        fl = (C.c_uint32 * max(1, len(fw)))(*[len(x) for x in fw])
        rl = (C.c_uint32 * max(1, len(rv)))(*[len(x) for x in rv])
        self._check(self._L.hx_set_primers(self._h, len(primers), _cstrs(fw), fl, _cstrs(rv), rl, mismatch))
        self.n_pairs = len(primers)
        self.pairs = [list(p) for p in primers]

    def run_batch(self, arena: np.ndarray, offsets: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """hx_run_batch on HOST buffers -> (n_records, n_pairs) structured array"""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if out is None:
            out = np.zeros((n, self.n_pairs), dtype=RESULT_DTYPE)
        ap = arena.ctypes.data if arena is not None and arena.size else None
        self._check(self._L.hx_run_batch(self._h, ap, offsets.ctypes.data, n, out.ctypes.data))
        return out

    def batch_stats(self) -> dict:
        """hx_batch_stats: what the last run_batch / run_batch_packed call moved"""
        v = (C.c_uint64 * 7)()
        self._check(self._L.hx_batch_stats(self._h, v, 7))
        return {"h2d_bytes": int(v[0]), "d2h_bytes": int(v[1]), "packed_text": bool(v[2]), "slabs": int(v[3]), "pack_threads": int(v[4]),
                "slabs_packed": int(v[5]), "mode": ("ascii", "packed_on_the_way", "mixed", "packed_arena")[int(v[6]) & 3]}

    def run_batch_packed(self, packed: np.ndarray, rec_flags: np.ndarray, offsets: np.ndarray,
                         out: np.ndarray | None = None) -> np.ndarray:
        """hx_run_batch_packed on a packed HOST arena (pack_records) -> (n_records, n_pairs) structured array"""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if out is None:
            out = np.zeros((n, self.n_pairs), dtype=RESULT_DTYPE)
        self._check(self._L.hx_run_batch_packed(self._h, packed.ctypes.data, rec_flags.ctypes.data if n else None,
                                                offsets.ctypes.data, n, out.ctypes.data))
        return out

    def run_batch_fasta(self, arena: np.ndarray, offsets: np.ndarray, ids: Sequence[str], tails: Sequence[str]):
        """hx_run_batch_fasta: (results, fasta_bytes) with the FASTA text of utils.rs:354-369 assembled on the GPU"""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros((n, self.n_pairs), dtype=RESULT_DTYPE)
        idb = [_b(i) for i in ids]
        id_off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            id_off[1:] = np.cumsum([len(x) for x in idb])
        tb = [_b(t) for t in tails]
        tail_off = np.zeros(len(tb) + 1, dtype=np.uint32)
        tail_off[1:] = np.cumsum([len(x) for x in tb])
        text, tlen = C.c_void_p(), C.c_uint64()
        ap = arena.ctypes.data if arena is not None and arena.size else None
        self._check(self._L.hx_run_batch_fasta(self._h, ap, offsets.ctypes.data, n, out.ctypes.data, b"".join(idb),
                                               id_off.ctypes.data, b"".join(tb), tail_off.ctypes.data, C.byref(text),
                                               C.byref(tlen)))
        return out, (C.string_at(text.value, tlen.value) if tlen.value else b"")

    def run_batch_text(self, arena: np.ndarray, offsets: np.ndarray, ids: Sequence[str], fa_tails: Sequence[str],
                       gff_tails: Sequence[str], stream=False):
        """hx_run_batch_text: (results, fasta_bytes, gff_bytes) — both output texts assembled on the GPU.
        stream=True (or a callable sink(kind, bytes) -> 0/1) goes through hx_run_batch_text_stream."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros((n, self.n_pairs), dtype=RESULT_DTYPE)
        idb = [_b(i) for i in ids]
        id_off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            id_off[1:] = np.cumsum([len(x) for x in idb])

        def blob(items):
            bs = [_b(t) for t in items]
            off = np.zeros(len(bs) + 1, dtype=np.uint32)
            off[1:] = np.cumsum([len(x) for x in bs])
            return b"".join(bs), off

        fb, fo = blob(fa_tails)
        gb, go = blob(gff_tails)
        ft, fl, gt, gl = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        ap = arena.ctypes.data if arena is not None and arena.size else None
        if stream:  # hx_run_batch_text_stream: the text arrives in chunks through a callback
            parts = ([], [])

            def sink(_user, kind, ptr, nbytes):
                parts[kind].append(C.string_at(ptr, nbytes))
                return int(stream(kind, parts[kind][-1])) if callable(stream) else 0

            cb = _lib.TEXT_SINK(sink)
            self._check(self._L.hx_run_batch_text_stream(self._h, ap, offsets.ctypes.data, n, out.ctypes.data,
                                                         b"".join(idb), id_off.ctypes.data, fb, fo.ctypes.data, gb,
                                                         go.ctypes.data, cb, None, C.byref(fl), C.byref(gl)))
            fa, gf = b"".join(parts[0]), b"".join(parts[1])
            assert len(fa) == fl.value and len(gf) == gl.value
            return out, fa, gf
        self._check(self._L.hx_run_batch_text(self._h, ap, offsets.ctypes.data, n, out.ctypes.data, b"".join(idb),
                                              id_off.ctypes.data, fb, fo.ctypes.data, gb, go.ctypes.data, C.byref(ft),
                                              C.byref(fl), C.byref(gt), C.byref(gl)))
        return (out, C.string_at(ft.value, fl.value) if fl.value else b"",
                C.string_at(gt.value, gl.value) if gl.value else b"")

    def run_batch_device(self, d_arena: int, d_offsets: int, n_records: int, arena_bytes: int, d_out: int,
                         stream: int = 0, dev_slot: int = 0):
        """hx_run_batch_device on raw device pointers (ints); asynchronous when stream != 0"""
        self._check(self._L.hx_run_batch_device(self._h, dev_slot, d_arena, d_offsets, n_records, arena_bytes, d_out,
                                                stream if stream else None))

    def find_all_end(self, pattern, text, k: int, fold_case: bool = False):
        """Myers::find_all_lazy(text, k).collect() on the GPU -> [(end, dist)]"""
        pattern, text = _b(pattern), _b(text)
        n = len(text)
        ends = np.empty(max(n, 1), dtype=np.uint64)
        dists = np.empty(max(n, 1), dtype=np.uint8)
        tb = np.frombuffer(text, dtype=np.uint8) if n else np.zeros(1, np.uint8)
        cnt = self._L.hx_find_all_end(self._h, pattern, len(pattern), tb.ctypes.data, n, k, 1 if fold_case else 0,
                                      ends.ctypes.data, dists.ctypes.data, n)
        if cnt < 0:
            raise HyperexError(int(cnt), self._L.hx_last_error(self._h).decode())
        return [(int(ends[i]), int(dists[i])) for i in range(cnt)]

    def group_info(self):
        """(n_groups, 32-bit automata per base, 64-bit automata per base) of the current primer set"""
        g, a, b = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._L.hx_group_info(self._h, C.byref(g), C.byref(a), C.byref(b)))
        return g.value, a.value, b.value

    def launch_count(self) -> int:
        return int(self._L.hx_launch_count(self._h))

    def kernel_time(self, reset: bool = True):
        ms, n = C.c_double(), C.c_uint64()
        self._check(self._L.hx_kernel_time(self._h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    FILE_STATS = ("wall_s", "bases", "records", "batches", "chunk_s", "parse_s", "device_s", "write_s", "log_s",
                  "fasta_bytes", "gff_bytes", "parsers", "lanes", "writers", "setup_s", "close_s")

    def file_stats(self) -> dict:
        """hx_file_stats: where the time of the last get_hypervar_regions call went (stage seconds are summed over
        the threads of a stage)"""
        v = (C.c_double * len(self.FILE_STATS))()
        self._check(self._L.hx_file_stats(self._h, v, len(self.FILE_STATS)))
        return dict(zip(self.FILE_STATS, (float(x) for x in v)))

    def get_hypervar_regions(self, file: str, primers: Sequence[Sequence[str]], prefix: str, mismatch: int,
                             log=None):
        """utils.rs:231-414: writes <prefix>.fa and <prefix>.gff; log(level, msg) gets warn!/error! lines (log=None:
        no callback at all, the result table then stays on the device)"""
        cb = _lib.LOG_FN((lambda lvl, msg, _u: log(lvl, msg.decode("utf-8", "replace"))) if log else 0)
        fw = _cstrs([p[0] for p in primers])
        rv = _cstrs([p[1] for p in primers])
        self._check(self._L.hx_get_hypervar_regions(self._h, _b(file), len(primers), fw, rv, _b(prefix), mismatch, cb,
                                                    None))


def get_hypervar_regions(file: str, primers, prefix: str, mismatch: int, devices=(0,), log=None):
    """Drop-in for utils::get_hypervar_regions(file, primers, prefix, mismatch) (src/utils.rs:231-236)."""
    with Context(devices) as ctx:
        ctx.get_hypervar_regions(file, primers, prefix, mismatch, log=log)


def plan_groups(primers: Sequence[Sequence[str]], mismatch: int, fast_small_k: bool = True, long_filter: int = 1,
                group_np_max: int = 0):
    """hx_plan_groups (host only, no GPU): (n_groups, 32-bit automata per base, 64-bit automata per base, group of
    every pair) for what hx_set_primers would build from this primer set and these options"""
    L = _lib.load()
    fw = [_b(p[0]) for p in primers]
    rv = [_b(p[1]) for p in primers]
    fl = (C.c_uint32 * max(1, len(fw)))(*[len(x) for x in fw])
    rl = (C.c_uint32 * max(1, len(rv)))(*[len(x) for x in rv])
    g, a, b = C.c_uint32(), C.c_uint32(), C.c_uint32()
    pg = np.zeros(max(1, len(primers)), np.uint32)
    rc = L.hx_plan_groups(len(primers), _cstrs(fw), fl, _cstrs(rv), rl, mismatch, int(fast_small_k), int(long_filter),
                          group_np_max, C.byref(g), C.byref(a), C.byref(b), pg.ctypes.data)
    if rc != 0:
        raise HyperexError(rc, (L.hx_last_error(None) or b"").decode())
    return g.value, a.value, b.value, pg[:len(primers)].tolist()


def plan_packing(primers: Sequence[Sequence[str]], mismatch: int, pair: int, role: int, alphabet: int = 0, byte: int = 65,
                 pack16: int = 1, pack8: int = 1):
    """hx_plan_packing (host only): dict(parts, index, m_scan, word) — how the small-k path packs this primer's automaton"""
    L = _lib.load()
    fw = [_b(p[0]) for p in primers]
    rv = [_b(p[1]) for p in primers]
    fl = (C.c_uint32 * max(1, len(fw)))(*[len(x) for x in fw])
    rl = (C.c_uint32 * max(1, len(rv)))(*[len(x) for x in rv])
    parts, idx, ms, w = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = L.hx_plan_packing(len(primers), _cstrs(fw), fl, _cstrs(rv), rl, mismatch, pack16, pack8, pair, role, alphabet, byte,
                           C.byref(parts), C.byref(idx), C.byref(ms), C.byref(w))
    if rc != 0:
        raise HyperexError(rc, (L.hx_last_error(None) or b"").decode())
    return {"parts": parts.value, "index": idx.value, "m_scan": ms.value, "word": w.value}


def plan_peq_word(primers: Sequence[Sequence[str]], mismatch: int, pair: int, role: int, alphabet: int, byte: int,
                  fast_small_k: bool = True, long_filter: int = 1, group_np_max: int = 0):
    """hx_plan_peq_word (host only): dict(scan, wm, verify, bits, m_scan, m_full) for one table entry of the plan"""
    L = _lib.load()
    fw = [_b(p[0]) for p in primers]
    rv = [_b(p[1]) for p in primers]
    fl = (C.c_uint32 * max(1, len(fw)))(*[len(x) for x in fw])
    rl = (C.c_uint32 * max(1, len(rv)))(*[len(x) for x in rv])
    w, ww, vw = C.c_uint64(), C.c_uint64(), C.c_uint64()
    bits, ms, mf = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = L.hx_plan_peq_word(len(primers), _cstrs(fw), fl, _cstrs(rv), rl, mismatch, int(fast_small_k), int(long_filter),
                            group_np_max, pair, role, alphabet, byte, C.byref(w), C.byref(ww), C.byref(vw), C.byref(bits),
                            C.byref(ms), C.byref(mf))
    if rc != 0:
        raise HyperexError(rc, (L.hx_last_error(None) or b"").decode())
    return {"scan": w.value, "wm": ww.value, "verify": vw.value, "bits": bits.value, "m_scan": ms.value, "m_full": mf.value}


def int_peak(variant: int, iters: int = 2000, device: int = 0) -> float:
    """measured integer throughput (1e9 lane-ops/s) of one instruction class; see hx_int_peak"""
    g = C.c_double()
    rc = _lib.load().hx_int_peak(device, variant, iters, C.byref(g))
    if rc != 0:
        raise HyperexError(rc, "hx_int_peak failed")
    return g.value


def bind_to_gpu_numa_node(device: int = 0):
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist), so that
    pinned arenas allocated afterwards are local to that GPU's PCIe root.  Host-side plumbing for the
    one-process-per-GPU layout; returns the cpu set or None when the topology is not exposed."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        return None
    return None


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8 numpy view of cudaHostAlloc'd memory (hx_alloc_pinned); keep the array alive while in use"""
    L = _lib.load()
    p = L.hx_alloc_pinned(max(nbytes, 1))
    if not p:
        raise MemoryError("hx_alloc_pinned failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED: dict = {}


def pinned_free(arr: np.ndarray):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        _lib.load().hx_free_pinned(p)
