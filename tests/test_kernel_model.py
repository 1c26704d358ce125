"""CPU statement of the device arithmetic of hx_myers_pair_kernel (hyperex_b200/csrc/hx_kernels.cu), checked against the
oracle: the claims DESIGN.md §3.1, §3.5 and §3.7 rest on are restated here in plain Python integers so that they are
pinned without a GPU —
  * top alignment: the pattern in the HIGH bits of a 32-bit word, unused low bits pv = 1 / mv = 0 / eq = 0, the score
    kept biased (dist - (k + 1)) so that "hit" is its sign (hx_step, utils.rs:305-309 through rust-bio's step);
  * case folding baked into the Peq rows (utils.rs:295) and the all-zero row of the null byte;
  * the Wu-Manber / Shift-Or automata of the small-k fast path, their initial state, and the distance of a hit taken as
    the NUMBER of levels whose top bit is set (the levels are monotone);
  * the suffix filter: candidates of the last 32 symbols, verified by a fresh 64-bit automaton over m + k bytes.
This is a model of the product's arithmetic, not the product: nothing here is imported by hyperex_b200/."""
import functools

import numpy as np
import pytest

from oracle import oracle as O

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


@functools.lru_cache(maxsize=None)
def _matches(sym: str, byte: int) -> bool:
    """utils.rs:251-268 through the oracle: does pattern symbol `sym` accept text byte `byte`?"""
    return bool(O.find_all_end(sym, bytes([byte]), 0)) if byte else False


def _peq_top(pat: str, byte: int, bits: int) -> int:
    """hx_api.cu peq_word + the row indexing of the kernel: the row of a raw text byte is that of its upper case"""
    u = byte - 32 if 97 <= byte <= 122 else byte
    m = len(pat)
    w = 0
    for i, c in enumerate(pat):
        if _matches(c, u):
            w |= 1 << (i + bits - m)
    return w


def _step(eq, pv, mv, s, mask):
    """hx_step of hx_kernels.cu on a `mask`-wide word"""
    top = (mask + 1) >> 1
    xv = eq | mv
    xh = ((((eq & pv) + pv) & mask) ^ pv) | eq
    ph = (mv | ~(xh | pv)) & mask
    mh = pv & xh
    s += (1 if ph & top else 0) - (1 if mh & top else 0)
    ph = (ph << 1) & mask
    mh = (mh << 1) & mask
    pv = (mh | ~(xv | ph)) & mask
    mv = ph & xv
    return pv, mv, s


def myers_top_aligned(pat: str, text: bytes, k: int, bits: int = 32):
    mask = M32 if bits == 32 else M64
    pv, mv, s = mask, 0, len(pat) - (k + 1)
    out = []
    for i, b in enumerate(text):
        pv, mv, s = _step(_peq_top(pat, b, bits), pv, mv, s, mask)
        if s < 0:
            out.append((i, s + k + 1))
    return out


def wu_manber_top_aligned(pat: str, text: bytes, k: int):
    """the KWM >= 0 branch of the scan kernel: Shift-Or levels R0..Rk, bit = 0 <=> prefix matches with <= d errors"""
    m = len(pat)
    pmask = (M32 << (32 - m)) & M32
    rv = [((M32 << (32 - m + d)) & M32) if d < m else 0 for d in range(k + 1)]  # level d: d lowest pattern bits clear
    out = []
    for i, b in enumerate(text):
        n = ~_peq_top(pat, b, 32) & pmask                                        # the complemented table (tables_wm)
        prev_old = rv[0]
        prev_old_s = (prev_old << 1) & M32
        prev_new = prev_old_s | n
        rv[0] = prev_new
        for d in range(1, k + 1):
            cur_old = rv[d]
            cur_old_s = (cur_old << 1) & M32
            cur_new = ((cur_old_s | n) & prev_old_s) & prev_old & ((prev_new << 1) & M32)
            rv[d] = cur_new
            prev_old, prev_old_s, prev_new = cur_old, cur_old_s, cur_new
        tops = [r >> 31 for r in rv]
        assert tops == sorted(tops, reverse=True), "levels must be monotone: <= d errors implies <= d + 1 errors"
        if not tops[k]:
            out.append((i, sum(tops)))                                           # hit path: one add per level
    return out


def _cases(seed, n, mmax):
    rng = np.random.default_rng(seed)
    letters = "ACGT" * 4 + "MRWSYKVHDBNU"
    for _ in range(n):
        m = int(rng.integers(1, mmax + 1))
        pat = "".join(letters[int(i)] for i in rng.integers(0, len(letters), m))
        txt = bytearray("".join("ACGTACGTNRUacgtn"[int(i)] for i in rng.integers(0, 16, int(rng.integers(0, 160)))).encode())
        if m <= len(txt) and rng.random() < 0.7:                                 # plant a (possibly mutated) site
            site = bytearray("".join(c if c in "ACGTU" else "A" for c in pat).encode())
            for _e in range(int(rng.integers(0, 3))):
                site[int(rng.integers(0, m))] = ord("ACGT"[int(rng.integers(0, 4))])
            p = int(rng.integers(0, len(txt) - m + 1))
            txt[p:p + m] = site
        yield pat, bytes(txt)


def test_top_aligned_biased_myers_equals_the_oracle():
    for pat, txt in _cases(1, 150, 32):
        for k in (0, 1, 3, len(pat)):
            assert myers_top_aligned(pat, txt, k) == O.find_all_end(pat, txt.upper(), k), (pat, txt, k)
    for pat, txt in _cases(2, 60, 64):                                           # 64-bit words (long_filter off, -m > 8)
        assert myers_top_aligned(pat, txt, 2, bits=64) == O.find_all_end(pat, txt.upper(), 2), (pat, txt)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_wu_manber_levels_equal_the_oracle(k):
    for pat, txt in _cases(10 + k, 150, 32):
        assert wu_manber_top_aligned(pat, txt, k) == O.find_all_end(pat, txt.upper(), k), (pat, txt, k)


def test_null_byte_row_leaves_both_automata_fresh():
    """the <= 15 bytes before a record start are cleared to 0: row 0 matches nothing and must not disturb a fresh state"""
    pat, txt = "GTGCCAGCMGCCGCGGTAA", b"ACGTGTGCCAGCAGCCGCGGTAATTT"
    pad = bytes(13)
    for k in (0, 2):
        ref = O.find_all_end(pat, txt, k)
        assert [(e - 13, d) for e, d in myers_top_aligned(pat, pad + txt, k)] == ref
        assert [(e - 13, d) for e, d in wu_manber_top_aligned(pat, pad + txt, k)] == ref


def test_suffix_filter_with_64_bit_verification_equals_the_oracle():
    """hx_verify_long as the kernel runs it: 32-bit scan of the last 32 symbols (either automaton), then a fresh
    top-aligned 64-bit automaton over the m + k bytes that end at the candidate"""
    rng = np.random.default_rng(21)
    for pat, txt in _cases(20, 80, 64):
        if len(pat) <= 32:
            continue
        m = len(pat)
        k = int(rng.integers(0, 3))
        got = []
        scan = wu_manber_top_aligned if rng.random() < 0.5 else myers_top_aligned
        for e, _d in scan(pat[-32:], txt, k):
            start = max(0, e - (m + k - 1))
            win = myers_top_aligned(pat, txt[start:e + 1], 64, bits=64)          # k = 64: report every position
            d = dict(win)[e - start]
            if d <= k:
                got.append((e, d))
        assert got == O.find_all_end(pat, txt.upper(), k), (pat, txt, k)


# ---- joining the windows of a long record: hx_combine_kernel / hx_seg_join -------------------
NONE = None


def _lexmin(a, b):
    return b if a is NONE or (b is not NONE and b < a) else a


def _segment(fh, rh, len_f, lo, hi):
    """what one scan thread leaves for its tile [lo, hi): bestF / bestR as (dist, end), pair best as (comb, f_end, r_end)"""
    bf = br = pair = NONE
    ev = sorted([(e, 1, d) for e, d in fh if lo <= e < hi] + [(e, 0, d) for e, d in rh if lo <= e < hi])
    for e, role, d in ev:                                  # reverse role first at equal positions (r_end > f_end)
        if role == 0:
            if bf is not NONE:
                cand = (bf[0] + d, bf[1], e)
                if pair is NONE or cand[0] < pair[0]:      # streaming rule: replace iff strictly smaller combined
                    pair = cand
            br = _lexmin(br, (d, e))
        elif e >= len_f - 1:                               # utils.rs:334-337
            bf = _lexmin(bf, (d, e))
    return {"bf": bf, "br": br, "pair": pair}


def _seg_join(L, R):
    """hx_seg_join: associative, order-sensitive"""
    pair = L["pair"]
    if L["bf"] is not NONE and R["br"] is not NONE:
        pair = _lexmin(pair, (L["bf"][0] + R["br"][0], L["bf"][1], R["br"][1]))
    pair = _lexmin(pair, R["pair"])
    return {"bf": _lexmin(L["bf"], R["bf"]), "br": _lexmin(L["br"], R["br"]), "pair": pair}


def _tree(segs):
    while len(segs) > 1:                                   # the shuffle tree of hx_combine_long_kernel: neighbours first
        segs = [_seg_join(segs[i], segs[i + 1]) if i + 1 < len(segs) else segs[i] for i in range(0, len(segs), 2)]
    return segs[0]


@pytest.mark.parametrize("seed", range(4))
def test_window_join_equals_the_literal_pairing_loop(seed):
    rng = np.random.default_rng(300 + seed)
    for _ in range(600):
        span = int(rng.integers(8, 200))
        f = sorted(set(int(x) for x in rng.integers(0, span, int(rng.integers(0, 12)))))
        r = sorted(set(int(x) for x in rng.integers(0, span, int(rng.integers(0, 12)))))
        fh = [(e, int(rng.integers(0, 4))) for e in f]
        rh = [(e, int(rng.integers(0, 4))) for e in r]
        len_f = int(rng.integers(1, 12))
        cuts = sorted(set([0, span] + [int(x) for x in rng.integers(1, span, int(rng.integers(0, 9)))]))
        segs = [_segment(fh, rh, len_f, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        want = O.pair_literal(fh, rh, len_f)               # (f_start, r_end, combined) or None
        for joined in (functools.reduce(_seg_join, segs), _tree(list(segs))):
            got = joined["pair"]
            got = None if got is NONE else (got[1] - (len_f - 1), got[2], got[0])
            assert got == want, (fh, rh, len_f, cuts)


# ---- hx_classify_kernel: SWAR search for 'U' / 'T' on 32-bit words (utils.rs:207-217) ---------------------------
def _swar_has(word: int, letter_lc: int) -> bool:
    lw = word | 0x20202020                                  # folds the case (utils.rs:210)
    x = lw ^ (letter_lc * 0x01010101)
    return bool(((x - 0x01010101) & M32) & ~x & 0x80808080)  # zero-byte test


def test_swar_letter_search_equals_bytewise():
    rng = np.random.default_rng(5)
    pool = np.frombuffer(b"ACGTUNacgtun\x00\xff\x55\x54\x75\x74\x35\x34\x15\x14uUtT", dtype=np.uint8)
    for _ in range(20000):
        bs = bytes(pool[rng.integers(0, len(pool), 4)]) if rng.random() < 0.8 else bytes(rng.integers(0, 256, 4, dtype=np.uint8))
        w = int.from_bytes(bs, "little")
        assert _swar_has(w, 0x75) == (b"U" in bs.upper() or b"u" in bs), bs
        assert _swar_has(w, 0x74) == (b"T" in bs.upper() or b"t" in bs), bs


# ---- hx_tile_build_kernel: a record as one tile or as equal 16-byte-granular windows with warm-up ----------------
def _tiles(length, single_max, tile_len, warm):
    if length == 0:
        return []
    if length <= single_max:
        return [(0, length, 0)]
    nt = (length + tile_len - 1) // tile_len
    tl = ((length + nt - 1) // nt + 15) & ~15
    nt = (length + tl - 1) // tl
    return [(i * tl, min(tl, length - i * tl), min(warm, i * tl)) for i in range(nt)]


def test_tiles_partition_the_record_and_windows_are_exact():
    rng = np.random.default_rng(9)
    for _ in range(3000):
        tile_len = int(rng.choice([16, 64, 256, 2048]))
        single_max = max(tile_len, int(rng.choice([16, 300, 4096])))
        warm = int(rng.integers(1, 80))
        length = int(rng.integers(0, 40000))
        ts = _tiles(length, single_max, tile_len, warm)
        assert sum(t[1] for t in ts) == length
        pos = 0
        for start, own, w in ts:
            assert start == pos and own > 0 and start % 16 == 0 and w == min(warm, start) and own <= max(single_max, tile_len + 15)
            pos += own
    # scanning every window from its warm-up start reproduces the whole-record hits (warm = m + k - 1)
    pat, k = "GTGCCAGCMGCCGCGGTAA", 2
    txt = ("ACGT" * 40 + "GTGCCAGCAGCCGCGGTAA" + "TTGA" * 33 + "GTGCCAGCCGCCGCGTAA" + "ACCA" * 21).encode()
    ref = O.find_all_end(pat, txt, k)
    for tile_len in (16, 48, 64):
        got = []
        for start, own, w in _tiles(len(txt), tile_len, tile_len, len(pat) + k - 1):
            got += [(e + start - w, d) for e, d in myers_top_aligned(pat, txt[start - w:start + own], k) if e - w >= 0]
        assert got == ref


# ---- the tables hx_set_primers uploads (host half, hx_plan_peq_word) against the model above ---------------------
def test_planned_tables_equal_the_model():
    import hyperex_b200 as hx
    rng = np.random.default_rng(31)
    letters = "ACGT" * 4 + "MRWSYKVHDBNU" + "anX"       # lower-case primer symbols never match (the text is folded, primers are not)
    text_bytes = [ord(c) for c in "ACGTUNRYKMSWBDHVacgtunx"] + [0, 1, 0x40, 0xFF]
    for trial in range(12):
        n = int(rng.integers(1, 20))
        prim = ["".join(letters[int(i)] for i in rng.integers(0, len(letters), int(rng.integers(1, 65)))) for _ in range(n + 1)]
        pairs = [[prim[int(rng.integers(0, len(prim)))], prim[int(rng.integers(0, len(prim)))]] for _ in range(n)]
        k = int(rng.choice([0, 2, 9]))
        lf = int(rng.choice([0, 2]))
        for _ in range(25):
            q, role, alpha = int(rng.integers(0, n)), int(rng.integers(0, 2)), int(rng.integers(0, 2))
            byte = int(text_bytes[int(rng.integers(0, len(text_bytes)))])
            e = hx.plan_peq_word(pairs, k, q, role, alpha, byte, long_filter=lf)
            full = pairs[q][0] if role == 0 else O.to_reverse_complement(pairs[q][1], O.RNA if alpha else O.DNA).decode()
            bits = e["bits"]
            assert e["m_full"] == len(full) and e["m_scan"] == min(len(full), bits)
            if lf == 0 or k > 8:
                assert bits == (64 if max(len(pairs[q][0]), len(pairs[q][1])) > 32 else 32)
            else:
                assert bits == 32
            scan = full[-bits:]
            want = _peq_top(scan, byte, bits)
            assert e["scan"] == want, (full, byte, bits)
            pmask = ((1 << len(scan)) - 1) << (bits - len(scan))
            assert e["wm"] == (~want & pmask)
            assert e["verify"] == (_peq_top(full, byte, 64) if len(full) > bits else 0)
    # every text byte of one long degenerate primer pair, both roles and alphabets
    pairs = [["ACGTUNRYKMSWBDHVanXACGTTGCAAGCTTGCATGCAAGCT", "TTGACANNRYACGU"]]
    for role in (0, 1):
        for alpha in (0, 1):
            full = pairs[0][0] if role == 0 else O.to_reverse_complement(pairs[0][1], O.RNA if alpha else O.DNA).decode()
            for byte in range(256):
                e = hx.plan_peq_word(pairs, 2, 0, role, alpha, byte, long_filter=2)
                assert e["scan"] == _peq_top(full[-32:], byte, 32) and e["m_full"] == len(full)
                assert e["verify"] == (_peq_top(full, byte, 64) if len(full) > 32 else 0)
