"""CPU suite for the product's host side: the C ABI loads and exports every declared symbol, and the
host helpers behave like the reference's (the cases mirror src/utils.rs:593-748).  No compute calls:
without a GPU the compute entry points must fail loudly, never fall back."""
import bz2
import ctypes
import gzip
import lzma
import os
import re

import numpy as np
import pytest

import hx_synth
import hyperex_b200 as hx
from hyperex_b200 import _lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hyperex_b200.h")).read()
    assert "hx_int_peak" not in header  # the microbenchmark is a bench-only export with its own header
    header += open(os.path.join(ROOT, "include", "hyperex_b200_bench.h")).read()
    declared = set(re.findall(r"\b(hx_[a-z_0-9]+)\s*\(", header)) - {"hx_log_fn"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hyperex_b200.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert lib.hx_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hx.HyperexError) as e:
        hx.Context((0,))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "hyperex_b200")
    for dp, _dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "_never_", f"{f} mentions the oracle"


# ---- mirrors of the reference unit tests (src/utils.rs:593-683) ------------------------------
def test_complement_dna():
    assert hx.to_complement("ATCGATCGATCGATCGRYKBVDHX", "dna") == "TAGCTAGCTAGCTAGCYRMVBHDX"


def test_complement_rna():
    assert hx.to_complement("AUCGAUCGAUCGAUCGRYKBVDHMXN", "rna") == "UAGCUAGCUAGCUAGCYRMVBHDKXN"


def test_reverse_complement():
    assert hx.to_reverse_complement("GTGCCAGCMGCCGCGGTAAN", "dna") == "NTTACCGCGGCKGCTGGCAC"


def test_sequence_type_dna_ok():
    assert hx.sequence_type("ATCGATCGATCG") == hx.Alphabet.Dna


def test_sequence_type_dna_ok2():
    assert hx.sequence_type("ATCGMTGCAATCG") == hx.Alphabet.Dna


def test_sequence_type_rna_ok():
    assert hx.sequence_type("AGCUUUGCA") == hx.Alphabet.Rna


def test_sequence_type_rna_ok2():
    assert hx.sequence_type("GUUUUAACCCAAM") == hx.Alphabet.Rna


def test_sequence_type_none():
    assert hx.sequence_type("ATCXXXRMGU") is None


def test_region_to_primer_ok():
    for r in O.supported_regions():
        assert hx.region_to_primer(r) == O.region_to_primer(r)
    assert hx.region_to_primer("v9v9") == []
    assert hx.supported_regions() == O.supported_regions() == hx.SUPPORTED_REGIONS
    assert hx.default_primers() == O.default_pairs()


def test_helpers_agree_with_oracle_on_random_strings():
    rng = np.random.default_rng(5)
    alpha = "ACGTUMRWSYKVHDBNXacgtn-"
    prim = O.default_pairs()
    for _ in range(300):
        s = "".join(alpha[i] for i in rng.integers(0, len(alpha), int(rng.integers(0, 40))))
        st = hx.sequence_type(s)
        assert (2 if st is None else st) == O.sequence_type(s)
        for name, a in (("dna", O.DNA), ("rna", O.RNA)):
            assert hx.to_complement(s, name).encode() == O.to_complement(s, a)
            assert hx.to_reverse_complement(s, name).encode() == O.to_reverse_complement(s, a)
    for f, _ in prim:
        for _, r in prim:
            assert hx.primers_to_region([f, r]) == O.primers_to_region(f, r)
    assert hx.primers_to_region(["AAA", "CCC"]) == ""


def test_file_to_vec(golden_dir, tmp_path):
    assert hx.file_to_vec(os.path.join(golden_dir, "primers.txt")) == [
        ["CCTACGGGNGGCWGCAG", "ATTACCGCGGCTGCTGG"], ["GTGCCAGCMGCCGCGGTAA", "GACTACHVGGGTATCTAATCC"]]
    bad = tmp_path / "bad.txt"
    bad.write_text("ACGT TTTT\n")
    with pytest.raises(ValueError, match="comma-separated"):  # utils.rs:141 / test_file_to_vec_no_ok
        hx.file_to_vec(str(bad))
    with pytest.raises(FileNotFoundError):
        hx.file_to_vec(str(tmp_path / "missing.txt"))
    crlf = tmp_path / "crlf.csv"
    crlf.write_bytes(b"AC,GT,extra\r\nTT,GG\r\n")
    assert hx.file_to_vec(str(crlf)) == [["AC", "GT"], ["TT", "GG"]]


# ---- FASTA ingest (rust-bio reader rules + niffler sniffing) ------------------------------------
def _read_all(path, max_bytes=0, threads=1):
    arenas, ids, lens = [], [], []
    for a, o, i in hx.read_fasta(path, max_bytes, threads=threads):
        arenas.append(a)
        ids += i
        lens += list(np.diff(o.astype(np.int64)))
    return (np.concatenate(arenas) if arenas else np.zeros(0, np.uint8)), ids, lens


def test_read_fasta_fixture_plain_and_gz(golden_dir):
    for name in ("test.fa", "test.fa.gz"):
        arena, ids, lens = _read_all(os.path.join(golden_dir, name))
        assert ids == ["Allorhizobium_borbori__DN316__EF125187"] and lens == [1353]
        oa, _oo, oi = O.parse_fasta(open(os.path.join(golden_dir, "test.fa"), "rb").read())
        assert arena.tobytes() == oa.tobytes() and ids == oi


@pytest.mark.parametrize("codec", ["plain", "gz", "bz2", "xz"])
def test_read_fasta_matches_oracle_reader(tmp_path, codec):
    a, o = hx_synth.gen_reads_mutated(40, seed=9, lower_frac=0.5, lmin=10, lmax=700)
    ids = [f"r{i}_x" if i % 3 else f"r{i} some description" for i in range(40)]
    buf = hx_synth.to_fasta(a, o, ids, width=60, crlf=True) + b">last\nAC GT\t\n\n>\n>never\nAA\n"
    path = tmp_path / f"in.fa.{codec}"
    path.write_bytes({"plain": lambda b: b, "gz": gzip.compress, "bz2": bz2.compress, "xz": lzma.compress}[codec](buf))
    oa, oo, oi = O.parse_fasta(buf)
    for mb, threads in ((0, 1), (1000, 1), (1000, 4), (0, 3)):  # whole file / many small batches; read-ahead threads
        arena, rid, lens = _read_all(str(path), mb, threads)
        assert rid == oi and lens == list(np.diff(oo.astype(np.int64))) and arena.tobytes() == oa.tobytes()


def test_read_fasta_fuzz_against_oracle_reader(tmp_path):
    """random line soup (CRLF, blank lines, bare '>', tabs, missing final newline, long lines that span the read
    buffer): the product reader and the oracle's restatement of rust-bio's reader give the same records"""
    rng = np.random.default_rng(2024)
    pieces = [b">", b">id", b">id desc more", b"> lead", b">\t", b"ACGT", b"acgtn", b"AC GT", b"", b" ", b"\t", b"NNNN",
              b">x>y", b"AC>GT", b"\r", b"ACGU\r", b">crlf id\r"]
    for case in range(300):
        n = int(rng.integers(0, 25))
        lines = []
        for _ in range(n):
            c = pieces[int(rng.integers(len(pieces)))]
            if rng.random() < 0.15:
                c = c + bytes(rng.choice(np.frombuffer(b"ACGTacgtNRYU -", np.uint8), int(rng.integers(1, 400))))
            lines.append(c)
        sep = b"\r\n" if rng.random() < 0.3 else b"\n"
        buf = sep.join(lines) + (sep if rng.random() < 0.7 else b"")
        if case % 50 == 0:  # a record whose single line is longer than the reader's 8 MB buffer
            buf = b">big\n" + b"ACGT" * (3 << 20) + b"\n" + buf
        if len(buf) < 5:    # niffler needs 5 bytes to sniff; shorter files are an open() error on both sides
            buf += b"\n\n\n\n\n"
        path = tmp_path / f"fuzz{case}.fa"
        path.write_bytes(buf)
        oa, oo, oi = O.parse_fasta(buf)
        for mb, threads in ((0, 1), (64, 1), (64, 3)):
            arena, rid, lens = _read_all(str(path), mb, threads)
            assert rid == oi, (case, buf[:200])
            assert lens == list(np.diff(oo.astype(np.int64))), (case, buf[:200])
            assert arena.tobytes() == oa.tobytes(), (case, buf[:200])
        path.unlink()


def test_read_fasta_errors(tmp_path):
    with pytest.raises(hx.HyperexError):
        list(hx.read_fasta(str(tmp_path / "missing.fa")))
    short = tmp_path / "short.fa"
    short.write_bytes(b">a\n")  # niffler: file too short to sniff
    with pytest.raises(hx.HyperexError):
        list(hx.read_fasta(str(short)))
    notfa = tmp_path / "not.fa"
    notfa.write_bytes(b"ACGTACGT\n>x\nAC\n")
    assert list(hx.read_fasta(str(notfa))) == []


def _batches(it):
    return [tuple(x.tobytes() if isinstance(x, np.ndarray) else x for x in b) for b in it]


@pytest.mark.parametrize("codec", ["plain", "gz"])
def test_read_ahead_reader_returns_the_single_thread_batches(tmp_path, codec):
    """hx_fasta_configure: the batches parsed ahead on background threads are, one by one and in order, the batches the
    calling thread alone cuts and parses for the same max_bytes — also when the iteration ends inside a later chunk
    (chunks parsed ahead of a malformed record are dropped) and when the caller stops early"""
    a, o = hx_synth.gen_reads_mutated(600, seed=5, lower_frac=0.2, lmin=1, lmax=900)
    ids = [f"r{i}" if i % 4 else f"r{i} desc {i}" for i in range(600)]
    good = hx_synth.to_fasta(a, o, ids, width=70)
    for name, buf in (("good", good), ("stops", good[: len(good) // 2] + b">\n>after\nACGT\n" + good[len(good) // 2:])):
        path = tmp_path / f"{name}.fa.{codec}"
        path.write_bytes(gzip.compress(buf) if codec == "gz" else buf)
        for mb in (5000, 40000, 0):
            one = _batches(hx.read_fasta(str(path), mb, threads=1))
            for threads in (2, 5, 0):
                assert _batches(hx.read_fasta(str(path), mb, threads=threads)) == one, (name, mb, threads)
        oa, oo, oi = O.parse_fasta(buf)
        assert b"".join(b[0] for b in one) == oa.tobytes() and sum((b[2] for b in one), []) == oi
    it = hx.read_fasta(str(path), 5000, threads=4)  # abandoned mid-way: close joins the threads
    next(it)
    it.close()


def test_packed_reader_equals_pack_records(tmp_path):
    """hx_fasta_next_packed = hx_pack_records of the batch hx_fasta_next returns: packed on the calling thread, by all
    the reader's threads, or ahead by the background threads — records with U, with T and U, with non-IUPAC bytes, empty
    records, odd and even batch lengths"""
    rng = np.random.default_rng(77)
    recs = []
    for i in range(500):
        n = int(rng.integers(0, 400))
        alpha = [b"ACGT", b"ACGU", b"ACGTU", b"ACGTRYSWKMBDHVNacgtn", b"ACGTXZ*-."][int(rng.integers(5))]
        recs.append(bytes(rng.choice(np.frombuffer(alpha, np.uint8), n)))
    buf = b"".join(b">s%d\n%s\n" % (i, r) for i, r in enumerate(recs) if i == 0 or r or i % 2)
    path = tmp_path / "p.fa"
    path.write_bytes(buf)
    for mb in (3000, 0):
        plain = list(hx.read_fasta(str(path), mb))
        for threads, ahead in ((1, False), (1, True), (4, False), (4, True)):
            got = list(hx.read_fasta_packed(str(path), mb, threads=threads, pack_ahead=ahead))
            assert len(got) == len(plain)
            for (pk, fl, off, ids, arena), (pa, po, pi) in zip(got, plain):
                assert arena.tobytes() == pa.tobytes() and off.tobytes() == po.tobytes() and ids == pi
                want_pk, want_fl = hx.pack_records(pa, po, 1)
                total = int(po[-1])
                assert fl.tobytes() == want_fl.tobytes(), (mb, threads, ahead)
                assert pk[: total // 2].tobytes() == want_pk[: total // 2].tobytes()
                if total & 1:
                    assert pk[total // 2] & 15 == want_pk[total // 2] & 15


# ---- pattern grouping (host half of hx_set_primers; hx_plan_groups needs no GPU) -------------------------
def test_builtin_regions_are_one_group_of_15_automata():
    pairs = hx.default_primers()
    for k in (0, 1, 3):
        assert hx.plan_groups(pairs, k)[:3] == (1, 15, 0)       # 20 searches -> 15 unique patterns, one launch
    # -m 2 on the Wu-Manber path: every built-in primer has a specific 16-symbol suffix, so two automata share a word
    # and the 15 automata stay one group (8 words x 3 levels of state) ...
    assert hx.plan_groups(pairs, 2)[:3] == (1, 15, 0)
    assert hx.plan_groups(pairs, 2, fast_small_k=False)[:3] == (1, 15, 0)
    # A pair with a primer whose last 16 symbols are too degenerate to filter with forms its own, unpacked class: it does
    # not unpack the others.
    vague = pairs + [["ACGTACGTAC" + "N" * 14, "TTGACATTGACAGGT"]]
    g, s32, s64, pg = hx.plan_groups(vague, 2)
    assert (g, s32, s64) == (2, 17, 0) and len(set(pg[:10])) == 1 and pg[10] != pg[0]
    # Unpacked automata at -m 2 run in groups of <= 8 (three bit-vectors each); primers shared between regions
    # (27F x3, 341F x2, 1492Rmod x3) stay together, so the split costs no duplicated automaton.
    g, s32, s64, pg = hx.plan_groups([[f + "N" * 16, r] for f, r in pairs], 2, long_filter=2)
    assert s32 == 15 and g == 2
    grp = dict(zip(hx.supported_regions(), pg))
    assert grp["v1v2"] == grp["v1v3"] == grp["v1v9"] == grp["v6v9"] == grp["v7v9"] and grp["v3v4"] == grp["v3v5"]


def test_small_k_pattern_packing_plan_and_table_words():
    """hx_plan_packing: which primers share a word on the small-k path, and the packed table words against a Python statement
    of DESIGN.md §3.8 (complemented Peq bits of the scanned suffix, top-aligned inside the part, unused low bits 0)"""
    pairs = hx.default_primers()
    for k, want in ((0, 4), (1, 2), (2, 2), (3, 1)):      # -m 0: four 8-symbol automata per word; -m 1..2: two 16-symbol ones
        assert {hx.plan_packing(pairs, k, q, role)["parts"] for q in range(len(pairs)) for role in (0, 1)} == {want}
    assert hx.plan_packing(pairs, 0, 0, 0, pack8=0)["parts"] == 2 and hx.plan_packing(pairs, 0, 0, 0, pack16=0)["parts"] == 1
    # a primer whose last 8 symbols are too degenerate (< 12 bits) keeps its whole group at two per word at -m 0 ...
    vague8 = [["ACGTACGTACGTAC" + "NNNNACGT", "TTGACATTGACAGGT"]]
    assert hx.plan_packing(vague8, 0, 0, 0)["parts"] == 2 and hx.plan_packing(vague8, 0, 0, 0, pack16=3)["parts"] == 4
    # ... and one whose last 16 are too degenerate is not packed at all unless forced
    vague16 = [["ACGTACGTAC" + "N" * 14, "TTGACATTGACAGGT"]]
    assert hx.plan_packing(vague16, 0, 0, 0)["parts"] == 1 and hx.plan_packing(vague16, 0, 0, 0, pack16=2)["parts"] == 2
    from oracle import oracle as O
    rng = np.random.default_rng(8)
    for k in (0, 1, 2):
        for q in (0, 3, 9):
            for role in (0, 1):
                for alphabet in (0, 1):
                    pat = pairs[q][0] if role == 0 else O.to_reverse_complement(pairs[q][1], O.RNA if alphabet else O.DNA).decode()
                    for byte in [ord(c) for c in "ACGTUNRYacgtn-"] + [0, 200, int(rng.integers(1, 256))]:
                        pl = hx.plan_packing(pairs, k, q, role, alphabet, byte)
                        w = 32 // pl["parts"]
                        suf = pat[-w:]
                        assert pl["m_scan"] == len(suf)
                        part = (pl["word"] >> (w * (pl["index"] % pl["parts"]))) & ((1 << w) - 1)
                        up = bytes([byte]).upper()    # utils.rs:295 (ASCII upper-casing)
                        model = 0
                        for i, c in enumerate(suf):   # bit (w - m + i) = 0 iff symbol i accepts the (upper-cased) text byte
                            if not (byte != 0 and O.find_all_end(c, up, 0)):   # rust-bio's ambig table, via the oracle
                                model |= 1 << (w - len(suf) + i)
                        assert part == model, (k, q, role, alphabet, byte, hex(part), hex(model))


def test_grouping_invariants_on_random_primer_sets():
    import hx_synth
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 90))
        prim = ["".join("ACGTRYN"[int(i)] for i in rng.integers(0, 7, int(rng.integers(1, 65)))) for _ in range(n + 2)]
        disjoint = trial % 2 == 0
        if disjoint:   # no primer shared between pairs: nothing ever has to be stepped twice
            pairs = [["A" + "ACGT"[i % 4] * (i // 4 + 1) + prim[i], "C" + "ACGT"[i % 4] * (i // 4 + 1) + prim[i + 1]][:2]
                     for i in range(min(n, 60))]
            pairs = [[f[:64], r[:64]] for f, r in pairs]
            n = len(pairs)
        else:          # primers drawn with replacement: large connected components that must be cut
            pairs = [[prim[int(rng.integers(0, len(prim)))], prim[int(rng.integers(0, len(prim)))]] for _ in range(n)]
        k = int(rng.choice([0, 2, 3, 9]))
        uniq = len({"f:" + p[0] for p in pairs} | {"r:" + p[1] for p in pairs})
        for fast, lf, cap in ((1, 2, 0), (0, 0, 0), (1, 2, 3), (0, 2, 2), (1, 1, 0)):
            g, s32, s64, pg = hx.plan_groups(pairs, k, fast_small_k=bool(fast), long_filter=lf, group_np_max=cap)
            assert len(pg) == n and max(pg) == g - 1 and set(pg) == set(range(g))
            assert 1 <= s32 + s64 <= 2 * n and g <= n
            if lf == 2 and k <= 8:
                assert s64 == 0                                   # forced suffix filter: no 64-bit automaton is stepped
            if disjoint and len({p[0] for p in pairs} & {hx.to_reverse_complement(p[1], "dna") for p in pairs}) == 0:
                assert s32 + s64 == uniq
    # 64-pair CSV of config 5: every pair placed, 123 automata with the filter, fewer launches than pairs
    pairs = hx_synth.gen_primer_pairs(54, seed=5, builtins=hx.default_primers())
    g, s32, s64, pg = hx.plan_groups(pairs, 2, fast_small_k=False)
    assert (s32, s64) == (123, 0) and g == 8
    g0, a0, b0, _ = hx.plan_groups(pairs, 2, fast_small_k=False, long_filter=0)
    assert b0 > 0 and a0 + b0 == 123
    # a primer whose last 32 symbols are too degenerate keeps its 64-bit automaton (every position would be a
    # candidate); a specific one is filtered; beyond -m 8 nothing is
    spec, vague = "ACGTTGCAAGCTTGCATGCAAGCTAGCTAGGATCCATGCA", "ACGTTGCA" + "N" * 32
    assert hx.plan_groups([[spec, "ACGTACGTACGTACGT"]], 2)[1:3] == (2, 0)
    assert hx.plan_groups([[vague, "ACGTACGTACGTACGT"]], 2)[1:3] == (0, 2)
    assert hx.plan_groups([[vague, "ACGTACGTACGTACGT"]], 2, long_filter=2)[1:3] == (2, 0)
    assert hx.plan_groups([[spec, "ACGTACGTACGTACGT"]], 9)[1:3] == (0, 2)
    assert hx.plan_groups([["ACGTTGCAAGCTTGCATGCAAGCTRGCYAGGATCCATGSW", "ACGTACGTACGTACGT"]], 8)[1:3] == (2, 0)


def test_plan_groups_rejects_what_set_primers_rejects():
    for bad in ([["", "ACGT"]], [["A" * 65, "ACGT"]], []):
        with pytest.raises(hx.HyperexError) as e:
            hx.plan_groups(bad, 0)
        assert e.value.code == -1


# ---- packed arena (hx_pack_records; pure host code) -----------------------------------------------------------
def _pack_model(arena, offsets):
    """numpy model of the packed layout: nibble i = code of base i, per-record evidence bits"""
    code = np.zeros(256, np.uint8)
    for c, ch in enumerate(hx.PACK_CODES):
        if c:
            code[ch] = c
            code[ch | 0x20] = c
    code[ord("U")] = code[ord("u")] = 4
    end = int(offsets[-1])
    full = np.zeros(end // 2 * 2 + 2, np.uint8)
    lo = int(offsets[0])
    full[lo:end] = code[arena[lo:end]]
    packed = (full[0::2] | (full[1::2] << 4)).astype(np.uint8)
    flags = np.zeros(len(offsets) - 1, np.uint8)
    up = arena.copy()
    low = (up >= ord("a")) & (up <= ord("z"))
    up[low] -= 32
    for r in range(len(offsets) - 1):
        seg = up[int(offsets[r]):int(offsets[r + 1])]
        f = (hx.REC_HAS_U if (seg == ord("U")).any() else 0) | (hx.REC_HAS_T if (seg == ord("T")).any() else 0)
        if ((code[seg] == 0)).any():
            f |= hx.REC_NON_IUPAC
        flags[r] = f
    return packed, flags


@pytest.mark.parametrize("isa", [0, 1, 2])       # scalar / AVX2 / AVX-512 VBMI (capped at what this CPU has)
@pytest.mark.parametrize("grain", [0, 192, 4096])  # bases per range the threads take: default (1 Mi) / finer than a record
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_pack_records_equals_the_model(threads, grain, isa, monkeypatch):
    monkeypatch.setenv("HYPEREX_PACK_ISA", str(isa))
    if grain:
        monkeypatch.setenv("HYPEREX_PACK_GRAIN", str(grain))
    rng = np.random.default_rng(77 + threads)
    alphabet = np.frombuffer(b"ACGT" * 12 + b"acgtnNRYSWKMBDHVUuXx-*.019 \t@[`{" + bytes([0, 1, 127, 128, 200, 255]), np.uint8)
    for trial in range(60):
        n = int(rng.integers(0, 50))
        lens = rng.integers(0, 3000 if trial % 3 else 40, n)
        if trial == 7:
            lens = np.array([0, 0, 1, 0, 1, 1, 0, 2, 0], dtype=np.int64)   # empty records at odd offsets
        if trial == 8:
            if grain == 192:
                continue
            lens = np.array([3_000_001, 5, 1_500_000], dtype=np.int64)    # several threads inside one record range
        if trial == 9:
            lens = np.array([700, 0, 0, 129, 127, 1, 0, 128, 256, 0, 63, 65, 640, 0], dtype=np.int64)  # ends on / next to block edges
        start = int(rng.integers(0, 5))
        offsets = np.zeros(len(lens) + 1, np.uint64)
        offsets[0] = start
        offsets[1:] = start + np.cumsum(lens)
        arena = alphabet[rng.integers(0, len(alphabet), int(offsets[-1]) + 3)]
        if trial % 4 == 0:  # clean reads: mostly ACGT, the vector path's common case
            arena = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(offsets[-1]) + 3)].copy()
        packed, flags = hx.pack_records(arena, offsets, threads)
        mp, mf = _pack_model(arena, offsets)
        lo, hi = int(offsets[0]), int(offsets[-1])
        # every nibble inside [offsets[0], offsets[-1]) is defined; compare nibble-wise
        got = np.stack([packed & 15, packed >> 4], 1).reshape(-1)[lo:hi]
        exp = np.stack([mp & 15, mp >> 4], 1).reshape(-1)[lo:hi]
        assert (got == exp).all(), trial
        assert (flags == mf).all(), trial
    # the evidence reproduces sequence_type (utils.rs:207-228)
    for s in ("ATCGATCGATCG", "ATCGMTGCAATCG", "AGCUUUGCA", "GUUUUAACCCAAM", "ATCXXXRMGU", "acgn", "ACGX", "", "T", "u"):
        a = np.frombuffer(s.encode(), np.uint8)
        _p, f = hx.pack_records(a, np.array([0, len(a)], np.uint64), 1)
        u, t, bad = bool(f[0] & hx.REC_HAS_U), bool(f[0] & hx.REC_HAS_T), bool(f[0] & hx.REC_NON_IUPAC)
        st = 1 if (u and not t) else 0 if (t and not u) else (2 if (u and t) else (2 if bad else 0))
        assert st == O.sequence_type(s), s
