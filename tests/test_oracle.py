"""CPU suite: pins the oracle (oracle/hx_oracle.c) against every known answer available for the path.

The reference's own tests assert no hit coordinate ("parity unpinned"), so the pins are
(1) the rust-bio 1.6.0 documentation vectors, (2) an independent O(mn) DP, (3) the reference unit-test
strings (src/utils.rs:593-683), (4) the config-1 golden hashes derived independently in SURVEY.md §8c.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import hx_synth
from oracle import oracle as O


# ---- (1) rust-bio doc / test vectors --------------------------------------
def test_rustbio_doc_vector_module_doc():
    assert O.find_all_end("TCCTAGGGC", "CGGTCCTGAGGGATTAGCAC", 2, hyperex_ambigs=False) == [(11, 2), (12, 2)]


def test_rustbio_doc_vector_find_all_end():
    assert O.find_all_end("TGAGCGT", "ACCGTGGATGAGCGCCATAG", 1, hyperex_ambigs=False) == [(13, 1), (14, 1)]


def test_rustbio_text_N_is_not_a_wildcard():
    assert min(d for _, d in O.find_all_end("TGAGCGT", "TGAGCNTA", 7, hyperex_ambigs=False)) == 1


def test_rustbio_ambig_is_asymmetric():
    assert min(d for _, d in O.find_all_end("TGRRCGTR", "TGABCNTR", 8, ambigs={"R": "AG"})) == 2
    assert min(d for _, d in O.find_all_end("TRANCGG", "GGATGNGCGCCATAG", 8, ambigs={"R": "AG", "N": "ACGT"})) == 2


# ---- (2) independent DP ---------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(6))
def test_bitparallel_equals_dp(seed):
    rng = np.random.default_rng(seed)
    pat_letters = "ACGTACGTACGTMRWSYKVHDBNU"
    txt_letters = "ACGTACGTACGTACGTNRYUX"
    for _ in range(150):
        m = int(rng.integers(1, 65))
        n = int(rng.integers(0, 200))
        k = int(rng.integers(0, min(m, 6) + 1)) if rng.random() < 0.9 else int(rng.integers(0, m + 3))
        pat = "".join(pat_letters[i] for i in rng.integers(0, len(pat_letters), m))
        txt = "".join(txt_letters[i] for i in rng.integers(0, len(txt_letters), n))
        if rng.random() < 0.5 and n > m:  # plant a mutated copy so that hits exist
            p = int(rng.integers(0, n - m + 1))
            site = hx_synth._resolve(rng, "".join(c if c in hx_synth.IUPAC else "A" for c in pat), 1)[0].tobytes()
            txt = txt[:p] + site.decode() + txt[p + m:]
        assert O.find_all_end(pat, txt, k) == O.dp_find_all_end(pat, txt, k), (pat, txt, k)


def test_bitparallel_equals_third_party_fuzzy_regex():
    """(end, dist) of the oracle against an implementation nobody here wrote: the `regex` module's fuzzy matching
    ((?:P){e<=d}, unit-cost insert / delete / substitute).  A hit (i, d) means: d is the smallest error count with which
    the pattern matches some substring of the text ENDING at i — checked as an anchored match of the reversed pattern on
    the reversed prefix; pattern symbols become character classes with the asymmetric table of utils.rs:251-263."""
    regex = pytest.importorskip("regex")
    amb = {"M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT", "K": "GT", "V": "ACGMRS", "H": "ACTMWY",
           "D": "AGTRWK", "B": "CGTSYK", "N": "ACGTMRWSYKVHDB"}
    rng = np.random.default_rng(5)
    for _ in range(250):
        m, n, k = int(rng.integers(2, 12)), int(rng.integers(5, 60)), int(rng.integers(0, 3))
        pat = "".join(rng.choice(list("ACGTRYNMWVB"), m))
        text = "".join(rng.choice(list("ACGTNRY"), n, p=[.22, .22, .22, .22, .04, .04, .04]))
        rp = "".join("[" + c + amb.get(c, "") + "]" for c in reversed(pat))
        res = [regex.compile("(?:%s){e<=%d}" % (rp, d)) for d in range(k + 1)]
        want = []
        for i in range(n):
            suffix = text[:i + 1][::-1]
            d = next((d for d in range(k + 1) if res[d].match(suffix)), None)
            if d is not None:
                want.append((i, d))
        assert O.find_all_end(pat, text, k) == want, (pat, text, k)


def test_window_restart_is_exact():
    """an automaton restarted m+k-1 bytes before j reports the same dist <= k hits"""
    rng = np.random.default_rng(7)
    for _ in range(200):
        m = int(rng.integers(2, 40))
        k = int(rng.integers(0, m + 2))
        pat = "".join("ACGTN"[i] for i in rng.integers(0, 5, m))
        txt = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
        full = O.find_all_end(pat, txt, k)
        cut = int(rng.integers(1, 299))
        start = max(0, cut - (m + k - 1))
        win = [(e + start, d) for e, d in O.find_all_end(pat, txt[start:], k) if e + start >= cut]
        assert win == [h for h in full if h[0] >= cut]


def test_suffix_filter_is_lossless_and_its_verification_exact():
    """hx_verify_long: the hits of a primer longer than 32 nt are a subset of the hits of its 32-symbol suffix
    (same k), and re-scanning the m + k bytes that end at a candidate gives the exact distance when it is <= k"""
    rng = np.random.default_rng(11)
    letters = "ACGT" * 5 + "MRWSYKVHDBN"
    for _ in range(150):
        m = int(rng.integers(33, 65))
        k = int(rng.integers(0, 9))
        pat = "".join(letters[i] for i in rng.integers(0, len(letters), m))
        txt = bytearray("".join("ACGT"[i] for i in rng.integers(0, 4, 400)).encode())
        for _s in range(3):  # planted, mutated sites (one at the very start)
            site = bytearray("".join(c if c in "ACGT" else "A" for c in pat).encode())
            for _e in range(int(rng.integers(0, k + 2))):
                site[int(rng.integers(0, len(site)))] = ord("ACGT"[int(rng.integers(0, 4))])
            p = 0 if _s == 0 else int(rng.integers(0, 400 - m))
            txt[p:p + m] = site
        txt = txt.decode()
        full = dict(O.find_all_end(pat, txt, k))
        cand = [e for e, _ in O.find_all_end(pat[-32:], txt, k)]
        assert set(full) <= set(cand)
        got = {}
        for e in cand:
            start = max(0, e - (m + k - 1))
            win = dict(O.find_all_end(pat, txt[start:e + 1], k))
            if e - start in win:
                got[e] = win[e - start]
        assert got == full, (pat, k)


def test_null_bytes_keep_a_fresh_automaton_fresh():
    """DESIGN: a fresh state stepped over never-matching bytes is unchanged (used to mask lead bytes)"""
    pat, txt = "GTGCCAGCMGCCGCGGTAA", "ACGTGTGCCAGCAGCCGCGGTAATTT"
    ref = O.find_all_end(pat, txt, 2)
    pad = "\x01" * 13
    assert [(e - 13, d) for e, d in O.find_all_end(pat, pad + txt, 2)] == ref


# ---- (3) reference unit-test strings (src/utils.rs:593-683) ---------------------------------
def test_complement_dna():
    assert O.to_complement("ATCGATCGATCGATCGRYKBVDHX", O.DNA) == b"TAGCTAGCTAGCTAGCYRMVBHDX"


def test_complement_rna():
    assert O.to_complement("AUCGAUCGAUCGAUCGRYKBVDHMXN", O.RNA) == b"UAGCUAGCUAGCUAGCYRMVBHDKXN"


def test_reverse_complement():
    assert O.to_reverse_complement("GTGCCAGCMGCCGCGGTAAN", O.DNA) == b"NTTACCGCGGCKGCTGGCAC"


def test_sequence_type():
    assert O.sequence_type("ATCGATCGATCG") == O.DNA
    assert O.sequence_type("ATCGMTGCAATCG") == O.DNA
    assert O.sequence_type("AGCUUUGCA") == O.RNA
    assert O.sequence_type("GUUUUAACCCAAM") == O.RNA
    assert O.sequence_type("ATCXXXRMGU") == O.NONE
    assert O.sequence_type("acgn") == O.DNA  # soft-masked, utils.rs:208-210
    assert O.sequence_type("") == O.DNA


def test_region_to_primer_ok():
    exp = {"v1v2": ["AGAGTTTGATCMTGGCTCAG", "ACTGCTGCSYCCCGTAGGAGTCT"],
           "v1v3": ["AGAGTTTGATCMTGGCTCAG", "ATTACCGCGGCTGCTGG"],
           "v1v9": ["AGAGTTTGATCMTGGCTCAG", "TACGGYTACCTTGTTAYGACTT"],
           "v3v4": ["CCTACGGGNGGCWGCAG", "GACTACHVGGGTATCTAATCC"],
           "v3v5": ["CCTACGGGNGGCWGCAG", "CCGTCAATTYMTTTRAGT"],
           "v4": ["GTGCCAGCMGCCGCGGTAA", "GGACTACHVGGGTWTCTAAT"],
           "v4v5": ["GTGYCAGCMGCCGCGGTAA", "CCCCGYCAATTCMTTTRAGT"],
           "v5v7": ["AACMGGATTAGATACCCKG", "ACGTCATCCCCACCTTCC"],
           "v6v9": ["TAAAACTYAAAKGAATTGACGGGG", "TACGGYTACCTTGTTAYGACTT"],
           "v7v9": ["YAACGAGCGCAACCC", "TACGGYTACCTTGTTAYGACTT"]}
    for r, pr in exp.items():
        assert O.region_to_primer(r) == pr
    assert O.region_to_primer("v2v3") == []
    assert O.supported_regions() == list(exp)
    assert O.primers_to_region(*exp["v4"]) == "v4"
    assert O.primers_to_region(*exp["v3v4"]) == "v3v4"
    assert O.primers_to_region("ACGT", "TTTT") == ""


# ---- (4) config 1 -----------------------------------------------------------------------------
def test_config1_golden(golden_dir):
    buf = open(os.path.join(golden_dir, "test.fa"), "rb").read()
    fa, gff, res = O.get_hypervar_regions(buf, O.default_pairs(), 0)
    # hashes derived independently in SURVEY.md §8c
    assert len(fa) == 2664 and hashlib.sha256(fa).hexdigest() == \
        "7c8b0083b64b8dbed0cafe64048e64ab1e9981c4f00b854e2ecb245a471d214c"
    assert len(gff) == 510 and hashlib.sha256(gff).hexdigest() == \
        "885ab68f83df0940714a9f9ca25ab01c1e23962f5bab05a54a5c918b348df043"
    assert fa == open(os.path.join(golden_dir, "config1_k0.fa"), "rb").read()
    regions = [(int(r["f_start"]) + 1, int(r["r_end"]) + 1) for r in res[0] if r["flags"] & 4]
    assert regions == [(268, 707), (268, 827), (417, 708), (417, 829), (683, 1100)]
    assert [int(r["flags"]) & 3 for r in res[0]] == [0, 2, 0, 3, 3, 3, 3, 3, 1, 1]


def test_config1_hit_table(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "golden.json")))
    text = O.parse_fasta(open(os.path.join(golden_dir, "test.fa"), "rb").read())[0].tobytes()
    ends_k0 = {"CCTACGGGNGGCWGCAG": 283, "GTGCCAGCMGCCGCGGTAA": 434, "GTGYCAGCMGCCGCGGTAA": 434,
               "AACMGGATTAGATACCCKG": 700, "TAAAACTYAAAKGAATTGACGGGG": 828, "YAACGAGCGCAACCC": 1019}
    for pat, end in ends_k0.items():
        assert O.find_all_end(pat, text, 0) == [(end, 0)]
    assert O.find_all_end("AGAGTTTGATCMTGGCTCAG", text, 0) == []
    for key, hits in meta["test_fa_hits"].items():
        name, k = key.split("|k")
        pat = name[2:] if name[0] == "F" else O.to_reverse_complement(name[2:], O.DNA).decode()
        assert O.find_all_end(pat, text, int(k)) == [tuple(h) for h in hits]


def test_gz_fixture_is_same_bytes(golden_dir):
    import gzip
    assert gzip.open(os.path.join(golden_dir, "test.fa.gz")).read() == \
        open(os.path.join(golden_dir, "test.fa"), "rb").read()


@pytest.mark.parametrize("name", ["c2_small", "c3_small", "mixed_small"])
def test_golden_batches(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    res = O.run_batch(z["arena"], z["offsets"], [list(p) for p in z["pairs"]], int(z["k"]), nthreads=3,
                      rebuild_per_record=True)
    assert res.tobytes() == z["res"].tobytes()


def test_synth_sites_hold_all_builtin_primers():
    a, o = hx_synth.gen_reads(32, seed=11)
    res = O.run_batch(a, o, O.default_pairs(), 0)
    assert (res["flags"] & 7 == 7).all()
    assert (res["combined"] == 0).all()


# ---- the pairing rule ---------------------------------------------------------------------------
def _stream_pair(f_hits, r_hits, len_f):
    """the list-free rule the CUDA kernel implements"""
    ev = sorted([(e, 0, d) for e, d in r_hits] + [(e, 1, d) for e, d in f_hits])  # reverse role first at equal pos
    best_f = None
    best = None
    for e, is_f, d in ev:
        if not is_f:
            if best_f is not None and e > best_f[1]:
                c = min(255, best_f[0] + d)
                if best is None or c < best[2]:
                    best = (best_f[1] - (len_f - 1), e, c)
        elif e >= len_f - 1 and (best_f is None or d < best_f[0]):
            best_f = (d, e)
    return best


@pytest.mark.parametrize("seed", range(4))
def test_streaming_pairing_equals_literal_loop(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(800):
        nf, nr = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        span = int(rng.integers(5, 60))
        f = sorted(set(int(x) for x in rng.integers(0, span, nf)))
        r = sorted(set(int(x) for x in rng.integers(0, span, nr)))
        fh = [(e, int(rng.integers(0, 4))) for e in f]
        rh = [(e, int(rng.integers(0, 4))) for e in r]
        len_f = int(rng.integers(1, 12))
        assert _stream_pair(fh, rh, len_f) == O.pair_literal(fh, rh, len_f)


# ---- FASTA reader rules (rust-bio fasta::Reader) --------------------------------
def test_parse_fasta_rules():
    buf = b">id1 desc here\r\nACGT\r\nAC GT  \r\n\r\n>id2\nTTTT\n>\n>id3\nGG\n"
    arena, offsets, ids = O.parse_fasta(buf)
    assert ids == ["id1", "id2"]  # the empty record ">" ends iteration
    assert arena.tobytes() == b"ACGTAC GTTTTT"
    assert list(offsets) == [0, 9, 13]
    assert O.parse_fasta(b"ACGT\n>x\nAA\n")[2] == []  # "Expected > at record start."
    assert O.parse_fasta(b">x\n")[2] == ["x"]
    assert O.parse_fasta(b"")[2] == []
