"""GPU parity suite (`-m gpu`): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs — bit-exact (integer work: hit coordinates, distances, flags, FASTA/GFF bytes)."""
import hashlib
import os

import numpy as np
import pytest

import hx_synth
import hyperex_b200 as hx
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _cmp(res, ref, arena=None, offsets=None, pairs=None, k=None):
    if res.tobytes() == ref.tobytes():
        return
    bad = np.argwhere((res["f_start"] != ref["f_start"]) | (res["r_end"] != ref["r_end"]) |
                      (res["flags"] != ref["flags"]) | (res["combined"] != ref["combined"]))
    r, p = bad[0]
    raise AssertionError(f"{len(bad)} mismatching (record, pair) cells; first at record {r} pair {p}: "
                         f"gpu={res[r, p]} oracle={ref[r, p]} k={k} pair={pairs[p] if pairs else None} "
                         f"len={int(offsets[r + 1] - offsets[r]) if offsets is not None else None}")


def _run(ctx, arena, offsets, pairs, k, **opts):
    for name in ("tile_len", "single_max"):
        ctx.set_option(name, opts.get(name, 0))
    ctx.set_option("slab_bytes", opts.get("slab_bytes", 128 << 20))
    ctx.set_option("text_tma", opts.get("text_tma", 0))
    ctx.set_option("fast_small_k", opts.get("fast", 1))
    ctx.set_option("long_filter", opts.get("lf", 1))
    # pattern packing of the small-k path: 0 off, 1 planner's choice, 2 two automata per word forced, 3 (-m 0) four forced;
    # pack8 = 0 keeps the planner from choosing four 8-symbol automata per word at -m 0
    ctx.set_option("pack16", opts.get("pack", 1))
    ctx.set_option("pack8", opts.get("pack8", 1))
    # 1: hx_run_batch packs every slab to 4 bits per base on its way to the device (ASCII in, nibble-text kernels)
    ctx.set_option("pack_h2d", opts.get("pack_h2d", -1))
    ctx.set_option("pack_threads", opts.get("pack_threads", 0))
    ctx.set_option("pack_piece_bytes", opts.get("pack_piece", 4 << 20))  # ring buffer size of that path
    ctx.set_primers(pairs, k)
    if opts.get("packed"):  # 4-bit packed arena (hx_pack_records + hx_run_batch_packed)
        pk, fl = hx.pack_records(arena, offsets, opts.get("pack_threads", 3))
        return ctx.run_batch_packed(pk, fl, offsets)
    return ctx.run_batch(arena, offsets)


# ---- find_all_lazy mirror ------------------------------------------------------------------------
def test_find_all_end_rustbio_vectors(ctx):
    assert ctx.find_all_end("TCCTAGGGC", "CGGTCCTGAGGGATTAGCAC", 2) == [(11, 2), (12, 2)]
    assert ctx.find_all_end("TGAGCGT", "ACCGTGGATGAGCGCCATAG", 1) == [(13, 1), (14, 1)]
    assert min(d for _, d in ctx.find_all_end("TGAGCGT", "TGAGCNTA", 7)) == 1


@pytest.mark.parametrize("tile_len", [0, 16, 48])
def test_find_all_end_random_vs_oracle(ctx, tile_len):
    rng = np.random.default_rng(1 + tile_len)
    ctx.set_option("tile_len", tile_len)
    try:
        for _ in range(120):
            m = int(rng.integers(1, 65))
            n = int(rng.integers(0, 600))
            k = int(rng.integers(0, min(m, 5) + 1)) if rng.random() < 0.85 else int(rng.integers(0, m + 3))
            pat = "".join("ACGTACGTMRWSYKVHDBNU"[i] for i in rng.integers(0, 20, m))
            txt = "".join("ACGTACGTACGTNRYUacgtn"[i] for i in rng.integers(0, 21, n))
            fold = bool(rng.integers(0, 2))
            ref = O.find_all_end(pat, txt.upper() if fold else txt, k)
            assert ctx.find_all_end(pat, txt, k, fold_case=fold) == ref, (pat, txt, k, fold)
    finally:
        ctx.set_option("tile_len", 0)


def test_find_all_end_capacity_is_reported_not_truncated_silently(ctx):
    import ctypes as C
    L = hx._lib.load()
    ends = np.zeros(4, np.uint64)
    dists = np.zeros(4, np.uint8)
    text = b"ACGT" * 50
    total = L.hx_find_all_end(ctx._h, b"AC", 2, text, len(text), 2, 0, ends.ctypes.data, dists.ctypes.data, 4)
    assert total == len(text)  # k >= m: every position is a hit
    assert list(ends) == [0, 1, 2, 3]


# ---- config 1: the reference's own CPU-runnable case ----------------------------------------------
@pytest.mark.parametrize("host_format", [0, 1])
@pytest.mark.parametrize("k", [0, 2])
def test_config1_bytes(ctx, golden_dir, tmp_path, k, host_format, monkeypatch):
    # host_format = 1: the host formatter (HYPEREX_HOST_FORMAT=1) instead of the on-device text assembly
    monkeypatch.setenv("HYPEREX_HOST_FORMAT", str(host_format))
    prefix = str(tmp_path / "hyperex_out")
    logs = []
    for f in ("test.fa", "test.fa.gz"):
        for ext in (".fa", ".gff"):
            if os.path.exists(prefix + ext):
                os.remove(prefix + ext)
        ctx.get_hypervar_regions(os.path.join(golden_dir, f), hx.default_primers(), prefix, k,
                                 log=lambda lvl, msg: logs.append((lvl, msg)))
        fa, gff = open(prefix + ".fa", "rb").read(), open(prefix + ".gff", "rb").read()
        assert fa == open(os.path.join(golden_dir, f"config1_k{k}.fa"), "rb").read()
        assert gff == open(os.path.join(golden_dir, f"config1_k{k}.gff"), "rb").read()
        assert hashlib.sha256(fa).hexdigest() == "7c8b0083b64b8dbed0cafe64048e64ab1e9981c4f00b854e2ecb245a471d214c"
        assert hashlib.sha256(gff).hexdigest() == "885ab68f83df0940714a9f9ca25ab01c1e23962f5bab05a54a5c918b348df043"
    if k == 0:
        msgs = [m for _, m in logs[:6]]
        assert msgs[0] == "Sequence Allorhizobium_borbori__DN316__EF125187 is short (1353 bp). Some regions may not be found."
        assert msgs[1] == "Both primers not found for region v1v2 in sequence Allorhizobium_borbori__DN316__EF125187"
        assert msgs[2] == ("Forward primer AGAGTTTGATCMTGGCTCAG not found for region v1v3 in sequence "
                           "Allorhizobium_borbori__DN316__EF125187")
        assert msgs[4] == ("Reverse primer TACGGYTACCTTGTTAYGACTT not found for region v6v9 in sequence "
                           "Allorhizobium_borbori__DN316__EF125187")


@pytest.mark.parametrize("text_tma", [0, 1])
@pytest.mark.parametrize("name", ["c2_small", "c3_small", "mixed_small"])
def test_golden_batches(ctx, golden_dir, name, text_tma):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    pairs = [list(p) for p in z["pairs"]]
    for tile in ((0, 0), (4096, 4096)):  # windows (auto for a small batch) and whole-record tiles
        for fast in (0, 1):               # Myers everywhere / Wu-Manber fast path where k <= 2
            res = _run(ctx, z["arena"], z["offsets"], pairs, int(z["k"]), tile_len=tile[0], single_max=tile[1],
                       text_tma=text_tma, fast=fast)
            _cmp(res, z["res"], z["arena"], z["offsets"], pairs, int(z["k"]))


# ---- seeded batches of the BASELINE.json shapes at oracle-friendly sizes -----------------------------
@pytest.mark.parametrize("fast,pack,pack8", [(0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)])
def test_c2_reads_all_regions_k0(ctx, fast, pack, pack8):
    """fast 0: Myers; fast 1 + pack 0: Shift-Or, one automaton per word; fast 1 + pack 1: suffix automata + verification of
    the rest of the primer — two 16-symbol automata per word (pack8 0) or four 8-symbol ones (library default)"""
    a, o = hx_synth.gen_reads(3000, seed=22)
    res = _run(ctx, a, o, hx.default_primers(), 0, fast=fast, pack=pack, pack8=pack8)
    if fast and pack:
        assert hx.plan_groups(hx.default_primers(), 0)[:3] == (1, 15, 0)
    _cmp(res, O.run_batch(a, o, O.default_pairs(), 0, nthreads=8), a, o, O.default_pairs(), 0)
    assert (res["flags"] & 7 == 7).all()


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_c3_degenerate_primers_both_strands(ctx, k, fast):
    a, o = hx_synth.gen_reads_mutated(1500, seed=30 + k, max_edits=3, revcomp_frac=0.5)
    pairs = [hx.region_to_primer("v3v4"), hx.region_to_primer("v4v5")]
    ref = O.run_batch(a, o, pairs, k, nthreads=8)
    for pack in ((1, 0) if fast else (1,)):
        _cmp(_run(ctx, a, o, pairs, k, fast=fast, pack=pack), ref, a, o, pairs, k)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_all_regions_mutated_small_k_fast_path(ctx, k):
    """15 automata per thread on the Wu-Manber path (k <= 2), soft-masked + RNA + both strands, with windows"""
    a, o = hx_synth.gen_reads_mutated(1200, seed=130 + k, max_edits=3, revcomp_frac=0.4, lower_frac=0.3, rna_frac=0.1)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, k, nthreads=8)
    for tile in ((4096, 4096), (128, 128)):
        for pack, pack8 in ((1, 1), (1, 0), (0, 1)):  # suffix automata + verification (library default: four per word at -m 0, else two) / one per word
            _cmp(_run(ctx, a, o, pairs, k, tile_len=tile[0], single_max=tile[1], fast=1, pack=pack, pack8=pack8), ref, a, o, pairs, k)


@pytest.mark.parametrize("text_tma", [0, 1])
@pytest.mark.parametrize("tile_len,single_max", [(0, 0), (256, 256), (64, 64), (1024, 1500)])
def test_c4_softmasked_contigs_windows(ctx, tile_len, single_max, text_tma):
    a, o = hx_synth.gen_contigs(3, seed=41, length=60_000, copies=(3, 6), n_runs=3)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 3, nthreads=8)
    _cmp(_run(ctx, a, o, pairs, 3, tile_len=tile_len, single_max=single_max, text_tma=text_tma), ref, a, o, pairs, 3)


def test_c4_megabase_contigs(ctx):
    """BASELINE.json config 4 at a size the oracle still finishes in seconds: 2 Mb contigs, ~1000 windows each,
    warp-cooperative tile build and the warp-parallel join of hx_combine_long_kernel"""
    a, o = hx_synth.gen_contigs(3, seed=43, length=2_000_000, copies=(4, 7), n_runs=4)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 3, nthreads=8)
    assert (ref["flags"] & 4).any()
    for fast, tma in ((0, 0), (0, 1)):
        _cmp(_run(ctx, a, o, pairs, 3, fast=fast, text_tma=tma), ref, a, o, pairs, 3)
    ref1 = O.run_batch(a, o, pairs, 1, nthreads=8)
    _cmp(_run(ctx, a, o, pairs, 1, fast=1), ref1, a, o, pairs, 1)


def test_c5_primer_csv_64_pairs(ctx):
    pairs = hx_synth.gen_primer_pairs(54, seed=5, builtins=O.default_pairs())
    assert len(pairs) == 64 and any(len(p[0]) > 32 or len(p[1]) > 32 for p in pairs)
    a, o = hx_synth.gen_reads_mutated(300, seed=51, max_edits=2, revcomp_frac=0.3)
    # plant some of the random primers so that the long-pattern (64-bit) path reports hits
    rng = np.random.default_rng(52)
    a = a.copy()
    for r in range(0, 300, 3):
        f, rv = pairs[10 + (r // 3) % 54]
        s = int(o[r])
        fs = hx_synth._resolve(rng, f, 1)[0]
        rs = hx_synth._resolve(rng, O.to_reverse_complement(rv, O.DNA).decode(), 1)[0]
        a[s + 100:s + 100 + len(fs)] = fs
        a[s + 700:s + 700 + len(rs)] = rs
    ref = O.run_batch(a, o, pairs, 2, nthreads=8)
    assert (ref[:, 10:]["flags"] & 4).any()
    for tma, lf in ((0, 1), (1, 1), (0, 0), (1, 0)):  # lf: suffix filter for primers > 32 nt / 64-bit groups
        _cmp(_run(ctx, a, o, pairs, 2, text_tma=tma, lf=lf), ref, a, o, pairs, 2)


@pytest.mark.parametrize("k", [0, 1, 2, 3, 6, 8, 9])
def test_long_primers_suffix_filter(ctx, k):
    """primers of 33..64 nt are scanned as the 32-bit automaton of their last 32 symbols and every candidate is
    verified against the whole primer (hx_verify_long): same table as the oracle and as the 64-bit groups, with
    sites planted at record starts / ends, mutated inside and outside the suffix, degenerate primers whose suffix
    matches almost anywhere, RNA records and windows"""
    rng = np.random.default_rng(700 + k)
    letters = "ACGT" * 6 + "MRWSYKVHDBN"
    def prim(n, alphabet=letters):
        return "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n))
    pairs = [[prim(40), prim(33)], [prim(64), prim(20)], [prim(18), prim(64)], [prim(33), prim(33)],
             [prim(45, "ACGT") + "N" * 19, prim(36)],                       # suffix of N: every position is a candidate
             ["N" * 30 + prim(10, "ACGT"), prim(50, "ACGTN")],
             hx.region_to_primer("v4")]
    pairs.append([pairs[0][0], pairs[1][1]])                                 # patterns shared between pairs
    recs = []
    for i in range(150):
        L = int(rng.integers(0, 700))
        seq = bytearray(hx_synth.random_acgt(rng, L).tobytes())
        for _s in range(int(rng.integers(0, 5))):
            f, r = pairs[int(rng.integers(0, len(pairs)))]
            site = f if rng.random() < 0.5 else O.to_reverse_complement(r, O.DNA).decode()
            sb = bytearray(hx_synth._resolve(rng, site, 1)[0].tobytes())
            sb = hx_synth._edit(rng, sb, int(rng.integers(0, min(k, 4) + 2)))
            if L >= len(sb):
                p = int(rng.choice([0, L - len(sb), int(rng.integers(0, L - len(sb) + 1))]))
                seq[p:p + len(sb)] = sb
        a1 = np.frombuffer(bytes(seq), dtype=np.uint8).copy()
        if i % 7 == 0:
            a1[a1 == ord("T")] = ord("U")
        if i % 5 == 0 and len(a1):
            lo, hi = sorted(rng.integers(0, len(a1) + 1, size=2))
            a1[lo:hi] |= 0x20
        recs.append(a1)
    a, o = hx_synth.pack(recs)
    ref = O.run_batch(a, o, pairs, k, nthreads=8)
    assert (ref["flags"] & 4).any()
    for tile in (0, 64, 4096):
        for fast, lf in ((1, 2), (0, 2), (1, 1), (0, 0)):  # lf 2: filter even the degenerate suffixes, 1: default, 0: off
            _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=fast, lf=lf), ref, a, o, pairs, k)
    if k <= 8:
        n32, n64 = ctx.group_info()[1:]
        ctx.set_option("long_filter", 2)
        ctx.set_primers(pairs, k)
        assert ctx.group_info()[2] == 0 and n64 > 0   # the filter leaves no 64-bit automaton to step


@pytest.mark.parametrize("k", [0, 1, 2])
def test_pattern_packing(ctx, k):
    """-m 0..2, two Shift-Or / Wu-Manber automata of <= 16 symbols per word: primers of 1..64 nt around the 16-symbol boundary, sites
    at record starts / ends, sites that match the 16-symbol suffix but not the rest of the primer (candidates the
    verification must reject), degenerate suffixes, odd and even pattern counts, RNA and soft-masked records, windows"""
    rng = np.random.default_rng(1616 + k)
    letters = "ACGT" * 6 + "MRWSYKVHDBN"
    def prim(n, alphabet=letters):
        return "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n))
    base = [[prim(15), prim(16)], [prim(17), prim(18)], [prim(24), prim(32)], [prim(33), prim(64)], [prim(1), prim(2)],
            [prim(8, "ACGT") + "N" * 16, prim(20)],          # suffix of N: every position is a candidate
            [prim(20, "ACGT"), "N" * 4 + prim(16, "ACGT")],
            hx.region_to_primer("v4"), hx.region_to_primer("v3v4")]
    for n_pairs in (1, 2, 5, len(base)):
        pairs = base[:n_pairs] if n_pairs > 2 else base[2:2 + n_pairs]
        recs = []
        for i in range(120):
            L = int(rng.integers(0, 600))
            seq = bytearray(hx_synth.random_acgt(rng, L).tobytes())
            for _s in range(int(rng.integers(0, 6))):
                f, r = pairs[int(rng.integers(0, len(pairs)))]
                site = f if rng.random() < 0.5 else O.to_reverse_complement(r, O.DNA).decode()
                sb = bytearray(hx_synth._resolve(rng, site, 1)[0].tobytes())
                sfx = 16 if rng.random() < 0.5 else 8           # damage the part in front of the 16- / 8-symbol suffix
                if rng.random() < 0.4 and len(sb) > sfx:
                    j = int(rng.integers(0, len(sb) - sfx))
                    sb[j] = ord("ACGT"[("ACGT".index(chr(sb[j])) + 1) % 4]) if chr(sb[j]) in "ACGT" else ord("A")
                elif rng.random() < 0.4:
                    sb = hx_synth._edit(rng, sb, int(rng.integers(1, k + 3)))
                if L >= len(sb):
                    p = int(rng.choice([0, L - len(sb), int(rng.integers(0, L - len(sb) + 1))]))
                    seq[p:p + len(sb)] = sb
            a1 = np.frombuffer(bytes(seq), dtype=np.uint8).copy()
            if i % 7 == 0:
                a1[a1 == ord("T")] = ord("U")
            if i % 5 == 0 and len(a1):
                lo, hi = sorted(rng.integers(0, len(a1) + 1, size=2))
                a1[lo:hi] |= 0x20
            recs.append(a1)
        a, o = hx_synth.pack(recs)
        ref = O.run_batch(a, o, pairs, k, nthreads=8)
        assert (ref["flags"] & 3).any() and (n_pairs < len(base) or (ref["flags"] & 4).any())
        for tile in (0, 64, 4096):
            for pack in (3, 2, 1, 0) if k == 0 else (2, 1, 0):  # 3: four 8-symbol automata per word forced (-m 0)
                _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=1, pack=pack), ref, a, o, pairs, k)
            _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=1, pack=2, packed=1), ref, a, o, pairs, k)
            if k == 0:
                _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=1, pack=3, packed=1), ref, a, o, pairs, k)


# ---- packed arena (north_star "2-bit/ASCII arena"): 4-bit text through hx_pack_records + hx_run_batch_packed --------
@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_packed_arena_equals_oracle(ctx, k):
    """soft-masked, RNA, both strands, garbage bytes (code 0), windows; every scan algorithm on the nibble text path"""
    a, o = hx_synth.gen_reads_mutated(1500, seed=440 + k, max_edits=3, revcomp_frac=0.4, lower_frac=0.3, rna_frac=0.15)
    a = a.copy()
    rng = np.random.default_rng(k)
    for r in rng.integers(0, 1500, 60):       # garbage inside records: 'X', '-', a digit, a high byte
        if o[r + 1] > o[r]:
            a[int(rng.integers(int(o[r]), int(o[r + 1])))] = int(rng.choice([ord("X"), ord("-"), ord("7"), 200]))
    for r in rng.integers(0, 1500, 40):       # 'U' in a DNA record / garbage in a record without T or U: alphabet None
        if o[r + 1] > o[r] + 10:
            a[int(o[r]) + 5] = ord("U") if r % 2 else ord("u")
    for r in rng.integers(0, 1500, 20):
        seg = a[int(o[r]):int(o[r + 1])]
        seg[(seg | 0x20) == ord("t")] = ord("A")
        seg[(seg | 0x20) == ord("u")] = ord("A")
        if len(seg):
            seg[len(seg) // 2] = ord("X") if r % 2 else ord("N")
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, k, nthreads=8)
    assert (ref["flags"] & 8).any() and (ref["flags"] & 16).any() and (ref["flags"] & 4).any()
    for tile in ((0, 0), (4096, 4096), (96, 96)):
        for fast in (0, 1):
            _cmp(_run(ctx, a, o, pairs, k, tile_len=tile[0], single_max=tile[1], fast=fast, packed=1), ref, a, o, pairs, k)
    _cmp(_run(ctx, a, o, pairs, k, packed=1, slab_bytes=50_000, pack_threads=8), ref, a, o, pairs, k)
    # ASCII in, packed on the way (pack_h2d): one slab, many slabs cut at unaligned / odd record starts
    # and slabs that go through the ring of piece buffers many times over
    for slab, th, piece in ((128 << 20, 0, 4 << 20), (50_000, 3, 4096), (1, 2, 4096), (333_333, 8, 8192), (128 << 20, 5, 4096),
                            (128 << 20, 0, 65536)):
        _cmp(_run(ctx, a, o, pairs, k, pack_h2d=1, slab_bytes=slab, pack_threads=th, pack_piece=piece), ref, a, o, pairs, k)
    # pack_h2d = 2: ASCII slabs from the front and packed slabs from the back of the batch in ONE call; which slab goes which
    # way depends on timing, the table must not
    for slab, th, piece in ((128 << 20, 0, 4 << 20), (50_000, 3, 4096), (1, 2, 4096), (333_333, 8, 8192), (20_000, 1, 4096),
                            (100_000, 16, 65536)):
        for _ in range(3):
            _cmp(_run(ctx, a, o, pairs, k, pack_h2d=2, slab_bytes=slab, pack_threads=th, pack_piece=piece), ref, a, o, pairs, k)
    st = ctx.batch_stats()
    assert st["slabs"] >= 1 and 0 <= st["slabs_packed"] <= st["slabs"]
    # an arena whose first record starts at an odd offset
    a2 = np.concatenate([np.frombuffer(b"GTGCCAG", np.uint8), a])
    o2 = o + np.uint64(7)
    _cmp(_run(ctx, a2, o2, pairs, k, packed=1, slab_bytes=30_000), ref, a2, o2, pairs, k)
    _cmp(_run(ctx, a2, o2, pairs, k, pack_h2d=1, slab_bytes=30_000), ref, a2, o2, pairs, k)
    _cmp(_run(ctx, a2, o2, pairs, k, pack_h2d=1, pack_piece=4096), ref, a2, o2, pairs, k)
    _cmp(_run(ctx, a2, o2, pairs, k, pack_h2d=2, slab_bytes=30_000), ref, a2, o2, pairs, k)
    _cmp(_run(ctx, a2, o2, pairs, k, pack_h2d=2, slab_bytes=64, pack_piece=4096, pack_threads=4), ref, a2, o2, pairs, k)


@pytest.mark.parametrize("k", [0, 2])
def test_packed_reader_feeds_run_batch_packed(ctx, tmp_path, k):
    """FASTA file -> hx_fasta_next_packed (chunks parsed and packed ahead by background threads) -> hx_run_batch_packed:
    the producer emits the packed arena; the table equals the oracle's over the oracle's own parse of the file"""
    a, o = hx_synth.gen_reads_mutated(2500, seed=610 + k, max_edits=2, revcomp_frac=0.3, lower_frac=0.3, rna_frac=0.2)
    ids = [f"r{i}" for i in range(2500)]
    buf = hx_synth.to_fasta(a, o, ids, width=80)
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    oa, oo, _oi = O.parse_fasta(buf)
    pairs = hx.default_primers()
    ref = O.run_batch(oa, oo, pairs, k, nthreads=8)
    assert (ref["flags"] & 4).any()
    _run(ctx, a[:int(o[1])], o[:2], pairs, k)  # options and primers of the suite's defaults
    for threads, ahead, mb in ((4, True, 200_000), (1, False, 0), (3, False, 1 << 20)):
        parts = []
        for pk, fl, off, _ids, ar in hx.read_fasta_packed(str(path), mb, threads=threads, pack_ahead=ahead, want_ids=False):
            got = ctx.run_batch_packed(pk, fl, off)
            assert got.tobytes() == ctx.run_batch(ar, off).tobytes()
            parts.append(got)
        _cmp(np.concatenate(parts), ref, oa, oo, pairs, k)


def test_pack_h2d_automatic_measures_both_modes(ctx):
    """pack_h2d = -1: the first six large batches alternate ASCII slabs and the mixed path (hx_batch_stats says which), the
    library then keeps one mode; every call returns the same table.  Small batches and hosts without packer threads stay ASCII."""
    a, o = hx_synth.gen_reads(48_000, seed=77)  # ~72 MB: above the 64 MB bar
    assert int(o[-1]) >= 64 << 20
    pairs = hx.default_primers()
    first = _run(ctx, a, o, pairs, 0, pack_h2d=-1, pack_threads=4)
    seen = [ctx.batch_stats()["mode"]]
    for _ in range(7):
        res = ctx.run_batch(a, o)
        seen.append(ctx.batch_stats()["mode"])
        assert res.tobytes() == first.tobytes()
    assert seen[:6] == ["ascii", "mixed"] * 3 and seen[6] == seen[7]
    _cmp(first[:300], O.run_batch(a[:int(o[300])], o[:301], pairs, 0, nthreads=8), a, o, pairs, 0)
    ctx.set_option("pack_h2d", -1)  # measures again
    ctx.run_batch(a, o)
    ctx.run_batch(a, o)
    st = ctx.batch_stats()
    assert st["mode"] == "mixed" and 0 <= st["slabs_packed"] <= st["slabs"]
    ctx.set_option("pack_threads", 1)  # fewer than 2 packer threads per device: ASCII
    for _ in range(2):
        ctx.run_batch(a, o)
        assert ctx.batch_stats()["mode"] == "ascii"
    ctx.set_option("pack_threads", 8)
    n = 2000  # a small batch never is
    for _ in range(2):
        ctx.run_batch(a[:int(o[n])], o[:n + 1])
        assert ctx.batch_stats()["mode"] == "ascii"
    ctx.set_option("pack_h2d", 1)
    ctx.run_batch(a, o)
    assert ctx.batch_stats()["mode"] == "packed_on_the_way" and ctx.batch_stats()["packed_text"]
    pk, fl = hx.pack_records(a, o)
    assert ctx.run_batch_packed(pk, fl, o).tobytes() == first.tobytes() and ctx.batch_stats()["mode"] == "packed_arena"


def test_packed_arena_contigs_long_primers_and_refusal(ctx):
    a, o = hx_synth.gen_contigs(3, seed=45, length=300_000)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 3, nthreads=8)
    _cmp(_run(ctx, a, o, pairs, 3, packed=1), ref, a, o, pairs, 3)
    _cmp(_run(ctx, a, o, pairs, 3, packed=1, tile_len=1024, single_max=1500), ref, a, o, pairs, 3)
    _cmp(_run(ctx, a, o, pairs, 3, pack_h2d=1, slab_bytes=200_000), ref, a, o, pairs, 3)
    _cmp(_run(ctx, a, o, pairs, 3, pack_h2d=1, pack_piece=16384), ref, a, o, pairs, 3)
    _cmp(_run(ctx, a, o, pairs, 3, pack_h2d=2, slab_bytes=400_000, pack_piece=16384), ref, a, o, pairs, 3)
    # primers of 33..64 nt: suffix filter + verification read the text as nibbles too; 64-bit groups
    a, o = hx_synth.gen_reads(300, seed=46)
    pairs = hx_synth.gen_primer_pairs(12, seed=5, builtins=hx.default_primers())
    assert max(len(x) for p in pairs for x in p) > 32
    for k, lf in ((2, 1), (2, 0), (0, 1), (9, 1)):
        ref = O.run_batch(a, o, pairs, k, nthreads=8)
        _cmp(_run(ctx, a, o, pairs, k, packed=1, lf=lf), ref, a, o, pairs, k)
    # a primer symbol outside ACGTURYSWKMBDHVN cannot be matched on packed text: refused, never approximated
    ctx.set_primers([["ACGTXACGT", "GGCC"]], 0)
    pk, fl = hx.pack_records(a, o)
    with pytest.raises(hx.HyperexError) as e:
        ctx.run_batch_packed(pk, fl, o)
    assert e.value.code == -4 and "packed" in str(e.value)
    ctx.set_primers([["acgtACGT", "GGCC"]], 0)   # lower-case primer symbols match nothing in either mode
    _cmp(ctx.run_batch_packed(pk, fl, o), O.run_batch(a, o, [["acgtACGT", "GGCC"]], 0), a, o)
    # ... and hx_run_batch with pack_h2d forced simply keeps the ASCII path for such a primer set
    a = a.copy()
    a[int(o[5]) + 3:int(o[5]) + 12] = np.frombuffer(b"ACGTXACGT", np.uint8)
    px = [["ACGTXACGT", "GGCC"], ["ACGTNACGT", "GGCC"]]
    _cmp(_run(ctx, a, o, px, 1, pack_h2d=1), O.run_batch(a, o, px, 1), a, o, px, 1)


# ---- edge cases the domain has -------------------------------------------------------------------------
def test_edge_records(ctx):
    recs = [b"", b"A", b"T", b"U", b"X", b"acgtn", b"ACGU" * 10 + b"T", b"N" * 50,
            b"GTGCCAGCMGCCGCGGTAA", b"GTGCCAGCAGCCGCGGTAA" + b"ATTAGAAACCCTTGTAGTCC",
            b"ccagcagccgcggtaat" + b"ACGT" * 30 + b"CCTACGGGAGGCAGCAG",  # reverse hit before forward hit
            b"CCTACGGGAGGCAGCAG", b"CTACGGGAGGCAGCAG" + b"A" * 40 + b"GGATTAGATACCCTGGTAGTC",
            b"CCUACGGGAGGCAGCAG" + b"ACGU" * 20 + b"GGAUUAGAUACCCUGGUAGUC",  # RNA record
            b""]
    a, o = hx_synth.pack(recs)
    for k in (0, 1, 3, 17, 24):
        pairs = hx.default_primers()
        ref = O.run_batch(a, o, pairs, k)
        for tma in (0, 1):
            _cmp(_run(ctx, a, o, pairs, k, text_tma=tma), ref, a, o, pairs, k)
            _cmp(_run(ctx, a, o, pairs, k, tile_len=16, single_max=16, text_tma=tma), ref, a, o, pairs, k)
        _cmp(_run(ctx, a, o, pairs, k, fast=0), ref, a, o, pairs, k)
        _cmp(_run(ctx, a, o, pairs, k, tile_len=16, single_max=16, fast=0), ref, a, o, pairs, k)
        for fast in (0, 1):
            _cmp(_run(ctx, a, o, pairs, k, packed=1, fast=fast), ref, a, o, pairs, k)
            _cmp(_run(ctx, a, o, pairs, k, tile_len=16, single_max=16, packed=1, fast=fast), ref, a, o, pairs, k)


def test_k_at_least_pattern_length_every_position_hits(ctx):
    # README example `-f ATCG -r GGCC -m 1` style: tiny primers, k close to / above m
    a, o = hx_synth.gen_reads_mutated(40, seed=61, lmin=5, lmax=300, sites=[])
    for pairs, k in (([["ATCG", "GGCC"]], 1), ([["ATCG", "GGCC"]], 4), ([["AC", "GTTGCA"]], 3),
                     ([["A", "C"], ["ACGTN", "T"]], 1), ([["A", "C"], ["AC", "T"]], 2), ([["ACG", "CA"]], 0)):
        ref = O.run_batch(a, o, pairs, k)
        for fast in (0, 1):
            _cmp(_run(ctx, a, o, pairs, k, tile_len=32, single_max=48, fast=fast), ref, a, o, pairs, k)


@pytest.mark.parametrize("seed", range(int(os.environ.get("HX_FUZZ_SEEDS", "40"))))
def test_random_primer_sets_fuzz(ctx, seed):
    """random IUPAC primer sets (1..64 nt, shared primers between pairs), random k, planted + mutated sites,
    mixed alphabets, windows of random size; both scan algorithms and both text paths"""
    rng = np.random.default_rng(1000 + int(os.environ.get("HX_FUZZ_BASE", "0")) + seed)
    letters = "ACGT" * 5 + "MRWSYKVHDBNU"
    n_pairs = int(rng.integers(1, 24))
    max_len = int(rng.choice([1, 6, 20, 31, 32, 33, 40, 64]))
    prim = ["".join(letters[int(i)] for i in rng.integers(0, len(letters), int(rng.integers(1, max_len + 1))))
            for _ in range(max(2, n_pairs))]
    pairs = [[prim[int(rng.integers(0, len(prim)))], prim[int(rng.integers(0, len(prim)))]] for _ in range(n_pairs)]
    k = int(rng.choice([0, 0, 1, 2, 3, 5, 9, 14]))
    recs = []
    for _ in range(int(rng.integers(20, 120))):
        L = int(rng.integers(0, 900))
        seq = bytearray(hx_synth.random_acgt(rng, L).tobytes())
        for _s in range(int(rng.integers(0, 6))):  # plant forward / reverse-complemented sites, some mutated
            f, r = pairs[int(rng.integers(0, n_pairs))]
            site = f if rng.random() < 0.5 else O.to_reverse_complement(r, O.DNA).decode()
            site = "".join(c if c in hx_synth.IUPAC else "A" for c in site)
            sb = bytearray(hx_synth._resolve(rng, site, 1)[0].tobytes())
            sb = hx_synth._edit(rng, sb, int(rng.integers(0, 3)))
            if L > len(sb):
                p = int(rng.integers(0, L - len(sb)))
                seq[p:p + len(sb)] = sb
        a1 = np.frombuffer(bytes(seq), dtype=np.uint8).copy()
        u = rng.random()
        if u < 0.15:
            a1[a1 == ord("T")] = ord("U")
        elif u < 0.3 and len(a1):
            a1[int(rng.integers(0, len(a1)))] = ord("X")
        if rng.random() < 0.3 and len(a1):
            lo, hi = sorted(rng.integers(0, len(a1) + 1, size=2))
            a1[lo:hi] |= 0x20
        recs.append(a1)
    a, o = hx_synth.pack(recs)
    ref = O.run_batch(a, o, pairs, k, nthreads=8)
    tile = int(rng.choice([0, 16, 64, 256, 4096]))
    for fast, tma, lf in ((0, 0, 1), (1, 0, 2), (0, 1, 1), (0, 0, 0)):
        _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=fast, text_tma=tma, lf=lf), ref, a, o, pairs,
             k)
    if k <= 2:  # small-k pattern packing: off / forced for every group (degenerate 16-symbol suffixes included)
        for pack in (0, 2) if k else (0, 2, 3):
            _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=1, pack=pack), ref, a, o, pairs, k)
    # the packed arena (4-bit text): Myers, the fast paths, suffix filters; slabs cut at unaligned record starts
    for fast, lf, pack in ((0, 1, 1), (1, 2, 2), (1, 1, 1), (0, 0, 0)) + (((1, 1, 3),) if k == 0 else ()):
        _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, fast=fast, lf=lf, pack=pack, packed=1,
                  slab_bytes=int(rng.choice([1, 3000, 128 << 20]))), ref, a, o, pairs, k)
    _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, pack_h2d=1, pack_threads=int(rng.integers(1, 9)),
              slab_bytes=int(rng.choice([1, 3000, 128 << 20])), pack_piece=int(rng.choice([4096, 8192, 4 << 20]))), ref, a, o,
         pairs, k)
    _cmp(_run(ctx, a, o, pairs, k, tile_len=tile, single_max=tile, pack_h2d=2, pack_threads=int(rng.integers(1, 9)),
              slab_bytes=int(rng.choice([1, 3000, 128 << 20])), pack_piece=int(rng.choice([4096, 8192, 4 << 20]))), ref, a, o,
         pairs, k)


def test_rna_records_use_rna_reverse_complement(ctx):
    a, o = hx_synth.gen_reads_mutated(200, seed=71, max_edits=1, revcomp_frac=0.0, rna_frac=0.6)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 2, nthreads=8)
    assert (ref["flags"] & 16).any() and not (ref["flags"] & 16).all()
    for tma in (0, 1):
        _cmp(_run(ctx, a, o, pairs, 2, text_tma=tma), ref, a, o, pairs, 2)


def test_primer_validation_errors(ctx):
    for bad in ([["", "ACGT"]], [["A" * 65, "ACGT"]], [["ACGT", "C" * 70]]):
        with pytest.raises(hx.HyperexError) as e:
            ctx.set_primers(bad, 0)
        assert e.value.code == -1
    ctx.set_primers([["A" * 64, "C" * 64]], 0)  # 64 is allowed (rust-bio: 0 < m <= 64)


def test_slab_pipeline_and_unaligned_offsets(ctx):
    a, o = hx_synth.gen_reads_mutated(700, seed=81, max_edits=2, lmin=1, lmax=900)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 2, nthreads=8)
    for slab in (1, 4096, 50_000, 1 << 20):
        _cmp(_run(ctx, a, o, pairs, 2, slab_bytes=slab), ref, a, o, pairs, 2)
        _cmp(_run(ctx, a, o, pairs, 2, slab_bytes=slab, pack_h2d=1), ref, a, o, pairs, 2)
        _cmp(_run(ctx, a, o, pairs, 2, slab_bytes=slab, pack_h2d=2), ref, a, o, pairs, 2)
    # an arena whose first record does not start at offset 0
    pad = 7
    a2 = np.concatenate([np.frombuffer(b"GTGCCAG", np.uint8), a])
    o2 = o + np.uint64(pad)
    _cmp(_run(ctx, a2, o2, pairs, 2, slab_bytes=30_000), ref, a2, o2, pairs, 2)


def test_device_resident_entry_point(ctx):
    import torch
    a, o = hx_synth.gen_reads(2000, seed=91)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 0, nthreads=8)
    ctx.set_option("tile_len", 0)
    ctx.set_option("single_max", 0)
    ctx.set_primers(pairs, 0)
    pad = (-len(a)) % 16 + 16
    d_arena = torch.from_numpy(np.concatenate([a, np.zeros(pad, np.uint8)])).cuda()
    d_off = torch.from_numpy(o.view(np.int64)).cuda()
    d_out = torch.zeros(len(o) - 1, len(pairs) * 12, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    ctx.run_batch_device(d_arena.data_ptr(), d_off.data_ptr(), len(o) - 1, len(a), d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    res = d_out.cpu().numpy().view(hx.RESULT_DTYPE).reshape(len(o) - 1, len(pairs))
    _cmp(res, ref, a, o, pairs, 0)


def _fa_tails(pairs):
    tails = []
    for f, r in pairs:
        lab = O.primers_to_region(f, r)
        tails.append((f" region={lab}" if lab else "") + f" forward={f} reverse={r}\n")
    return tails


@pytest.mark.parametrize("k,tile", [(0, 0), (2, 0), (3, 128)])
def test_on_device_fasta_extraction(ctx, k, tile):
    """the FASTA text assembled on the GPU equals the oracle's bytes (original case, output order)"""
    a, o = hx_synth.gen_reads_mutated(900, seed=140 + k, max_edits=2, revcomp_frac=0.3, lower_frac=0.5, rna_frac=0.1,
                                      lmin=1, lmax=1200)
    ids = [f"id{i}" if i % 3 else f"a_much_longer_identifier_{i}" for i in range(900)]
    pairs = hx.default_primers() + [["ACGTACGTAC", "TTGACA"]]
    for name in ("tile_len", "single_max"):
        ctx.set_option(name, tile)
    ctx.set_primers(pairs, k)
    res, text = ctx.run_batch_fasta(a, o, ids, _fa_tails(pairs))
    ref = O.run_batch(a, o, pairs, k, nthreads=8)
    _cmp(res, ref, a, o, pairs, k)
    fa, gff = O.format_outputs(a, o, ids, pairs, ref)
    assert text == fa
    gtails = []
    for f, r in pairs:
        lab = O.primers_to_region(f, r)
        gtails.append("\t.\t.\t.\tNote=" + (f"Hypervariable region {lab}" if lab else f"forward={f} reverse={r}") + "\n")
    res3, fa3, gff3 = ctx.run_batch_text(a, o, ids, _fa_tails(pairs), gtails)
    assert res3.tobytes() == ref.tobytes() and fa3 == fa and b"##gff-version 3\n" + gff3 == gff
    # a batch without any region gives an empty text
    a2, o2 = hx_synth.pack([b"", b"T", b"XX"])
    res2, text2 = ctx.run_batch_fasta(a2, o2, ["x", "y", "z"], _fa_tails(pairs))
    assert text2 == b"" and not (res2["flags"] & 4).any()


def test_text_stream_chunks(ctx):
    """hx_run_batch_text_stream: > 16 MiB of text arrives in order through the sink and equals the buffered call"""
    a, o = hx_synth.gen_reads(9000, seed=77)
    ids = [f"read_{i}" for i in range(9000)]
    pairs = hx.default_primers()
    for name in ("tile_len", "single_max"):
        ctx.set_option(name, 0)
    ctx.set_primers(pairs, 0)
    gtails = ["\t.\t.\t.\tNote=Hypervariable region " + O.primers_to_region(f, r) + "\n" for f, r in pairs]
    res, fa, gff = ctx.run_batch_text(a, o, ids, _fa_tails(pairs), gtails)
    chunks = []
    res2, fa2, gff2 = ctx.run_batch_text(a, o, ids, _fa_tails(pairs), gtails,
                                         stream=lambda kind, b: chunks.append((kind, len(b))) or 0)
    assert len(fa) > (32 << 20) and fa2 == fa and gff2 == gff and res2.tobytes() == res.tobytes()
    assert len(chunks) >= 4 and all(n <= (16 << 20) for _, n in chunks)
    assert [k for k, _ in chunks] == sorted(k for k, _ in chunks)          # FASTA chunks first, then GFF3
    ref = O.run_batch(a, o, pairs, 0, nthreads=8)
    rfa, rgff = O.format_outputs(a, o, ids, pairs, ref)
    assert fa2 == rfa and b"##gff-version 3\n" + gff2 == rgff
    with pytest.raises(hx.HyperexError):                                   # a failing sink is an error, not a truncation
        ctx.run_batch_text(a, o, ids, _fa_tails(pairs), gtails, stream=lambda kind, b: 1)
    res3, fa3, _ = ctx.run_batch_text(a, o, ids, _fa_tails(pairs), gtails, stream=True)   # context still usable
    assert fa3 == fa and res3.tobytes() == res.tobytes()


# ---- file-level seam: FASTA in, FASTA + GFF3 bytes out ---------------------------------------------------
@pytest.mark.parametrize("host_format,chunk", [(0, 0), (1, 0), (0, 20000), (1, 20000)])
@pytest.mark.parametrize("k,codec", [(0, "plain"), (2, "gz"), (3, "xz")])
def test_get_hypervar_regions_bytes(ctx, tmp_path, k, codec, host_format, chunk, monkeypatch):
    """both formatters (on-device text / HYPEREX_HOST_FORMAT=1), as one batch and as ~30 batches that finish out of
    order on the context's text lanes"""
    import gzip
    import lzma
    monkeypatch.setenv("HYPEREX_HOST_FORMAT", str(host_format))
    if chunk:
        monkeypatch.setenv("HYPEREX_CHUNK_BYTES", str(chunk))
    a, o = hx_synth.gen_reads_mutated(400, seed=95 + k, max_edits=2, revcomp_frac=0.3, lower_frac=0.4, rna_frac=0.1)
    ids = [f"seq{i}" if i % 2 else f"seq{i} len={int(o[i + 1] - o[i])}" for i in range(400)]
    buf = hx_synth.to_fasta(a, o, ids, width=70 if k else 0, crlf=bool(k == 2))
    path = tmp_path / "in.fa"
    path.write_bytes({"plain": lambda b: b, "gz": gzip.compress, "xz": lzma.compress}[codec](buf))
    pairs = hx.default_primers() + [["ACGTACGTAC", "TTGACA"]]  # one pair without a region label
    prefix = str(tmp_path / "out")
    ctx.set_option("tile_len", 0)
    ctx.set_option("single_max", 0)
    logs = []
    ctx.get_hypervar_regions(str(path), pairs, prefix, k, log=lambda lvl, msg: logs.append((lvl, msg)))
    fa, gff, res = O.get_hypervar_regions(buf, pairs, k)
    assert open(prefix + ".fa", "rb").read() == fa
    assert open(prefix + ".gff", "rb").read() == gff
    st = ctx.file_stats()
    assert st["records"] == 400 and st["fasta_bytes"] == len(fa) and st["gff_bytes"] == len(gff)
    assert (st["batches"] > 10) == bool(chunk) or codec != "plain"
    # the reference's warn!/error! lines arrive in record x pair order whatever lane ran the batch
    n_expected = int((res["flags"][:, 0] & 8).astype(bool).sum())                    # error: unrecognized type
    ok = ~(res["flags"][:, 0] & 8).astype(bool)
    n_expected += int((np.diff(o.astype(np.int64))[ok] <= 1500).sum())               # warn: short sequence
    n_expected += int(((res["flags"][ok] & 7) != 7).sum())                           # warn: one of the 4 not-found texts
    assert len(logs) == n_expected
    seen = [m.rsplit(" ", 1)[-1] for _, m in logs if " in sequence " in m and m.endswith(tuple("0123456789"))]
    order = [int(x[3:]) for x in seen if x.startswith("seq")]
    assert order == sorted(order)
    # without a log callback the result table never leaves the device
    ctx.get_hypervar_regions(str(path), pairs, prefix + "_nolog", k)
    assert open(prefix + "_nolog.fa", "rb").read() == fa and open(prefix + "_nolog.gff", "rb").read() == gff


# ---- several GPUs in one context (hx_create(dev_ids, n)): the shipped `--gpus N` path -------------------------
def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_gpus() < 2", reason="needs two GPUs")
@pytest.mark.parametrize("host_format", [0, 1])
def test_multi_device_context_equals_single_device(ctx, tmp_path, host_format, monkeypatch):
    monkeypatch.setenv("HYPEREX_HOST_FORMAT", str(host_format))
    monkeypatch.setenv("HYPEREX_CHUNK_BYTES", "30000")
    nd = min(_n_gpus(), 8)
    a, o = hx_synth.gen_reads_mutated(3000, seed=31, max_edits=2, revcomp_frac=0.3, lower_frac=0.2, rna_frac=0.1)
    pairs = hx.default_primers()
    ref = O.run_batch(a, o, pairs, 2, nthreads=8)
    buf = hx_synth.to_fasta(a, o, [f"seq{i}" for i in range(3000)], width=80)
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    fa, gff, _ = O.get_hypervar_regions(buf, pairs, 2)
    with hx.Context(tuple(range(nd))) as mctx:
        mctx.set_option("slab_bytes", 1 << 20)  # many slabs so every device gets work
        mctx.set_primers(pairs, 2)
        assert mctx.run_batch(a, o).tobytes() == ref.tobytes()
        for mode, slab, threads in ((1, 1 << 20, 4), (2, 1 << 20, 4), (2, 200_000, 2), (2, 64 << 20, 8)):
            # packed on the way / mixed slabs, spread over the devices of the context
            mctx.set_option("pack_h2d", mode)
            mctx.set_option("slab_bytes", slab)
            mctx.set_option("pack_threads", threads)
            assert mctx.run_batch(a, o).tobytes() == ref.tobytes(), (mode, slab, threads)
        mctx.set_option("pack_h2d", -1)
        logs = []
        mctx.get_hypervar_regions(str(path), pairs, str(tmp_path / "m"), 2, log=lambda lvl, msg: logs.append(msg))
        assert (tmp_path / "m.fa").read_bytes() == fa and (tmp_path / "m.gff").read_bytes() == gff
        assert mctx.file_stats()["lanes"] == 2 * nd and mctx.file_stats()["batches"] > 2 * nd
    logs1 = []
    ctx.get_hypervar_regions(str(path), pairs, str(tmp_path / "s"), 2, log=lambda lvl, msg: logs1.append(msg))
    assert logs == logs1 and (tmp_path / "s.fa").read_bytes() == fa


# ---- full-size properties (BASELINE.json config 2 shape; the oracle is too slow at this size) -------------
def test_full_size_properties(ctx):
    n = 200_000
    a, o = hx_synth.gen_reads(n, seed=2)
    pairs = hx.default_primers()
    res = _run(ctx, a, o, pairs, 0)
    # every planted site is found exactly, in gene order, inside the record
    assert (res["flags"] == 7).all() and (res["combined"] == 0).all()
    lens = np.diff(o.astype(np.int64))
    assert (res["r_end"] < lens[:, None]).all() and (res["f_start"] < res["r_end"]).all()
    # idempotence / order independence: a permuted batch gives the permuted table
    perm = np.random.default_rng(3).permutation(n)[:20000]
    recs = [a[int(o[i]):int(o[i + 1])] for i in perm]
    a2, o2 = hx_synth.pack(recs)
    res2 = _run(ctx, a2, o2, pairs, 0, slab_bytes=8 << 20)
    assert res2.tobytes() == res[perm].tobytes()
    # and a sample agrees with the oracle
    samp = np.arange(0, n, 97)
    a3, o3 = hx_synth.pack([a[int(o[i]):int(o[i + 1])] for i in samp])
    assert O.run_batch(a3, o3, pairs, 0, nthreads=8).tobytes() == res[samp].tobytes()
