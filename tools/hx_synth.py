"""Seeded synthetic inputs of the shapes BASELINE.json names.  Test / bench
infrastructure only (numpy); not part of the product and not the oracle."""
from __future__ import annotations

import numpy as np

IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "U": "U", "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT",
         "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "N": "ACGT"}

# Combined primer sites of a 16S rRNA gene, in gene order, at E. coli coordinates.  Each string
# contains every built-in forward primer / reverse-complemented reverse primer that overlaps that
# site (tests/test_oracle.py::test_synth_sites_hold_all_builtin_primers checks it at k = 0).
SITES_16S = [
    (8, "AGAGTTTGATCMTGGCTCAG"),          # 27F
    (331, "AGACTCCTACGGGRGGCAGCAGT"),     # rc(336R) + 341F
    (515, "GTGCCAGCAGCCGCGGTAAT"),        # 515F, 515F-Y, rc(534R)
    (781, "AACMGGATTAGATACCCKGGTAGTCC"),  # 799F, rc(805R), rc(806R)
    (919, "TAAAACTYAAAKGAATTGACGGGG"),    # 928F, rc(926Rb), rc(909-928R)
    (1100, "YAACGAGCGCAACCC"),            # 1100F
    (1176, "GGAAGGTGGGGATGACGT"),         # rc(1193R)
    (1492, "AAGTCRTAACAAGGTARCCGTA"),     # rc(1492Rmod)
]
ECOLI_LEN = 1542.0
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgtRYKMBVDHrykmbvdh", b"TGCAtgcaYRMKVBHDyrmkvbhd"):
    _COMP[a] = b


def revcomp_bytes(x: np.ndarray) -> np.ndarray:
    return _COMP[x[::-1]]


def random_acgt(rng: np.random.Generator, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def _resolve(rng, site: str, n: int) -> np.ndarray:
    """n x len(site) uint8 matrix with IUPAC codes resolved uniformly at random"""
    out = np.empty((n, len(site)), dtype=np.uint8)
    for j, ch in enumerate(site):
        alts = np.frombuffer(IUPAC[ch].encode(), dtype=np.uint8)
        out[:, j] = alts[rng.integers(0, len(alts), size=n)] if len(alts) > 1 else alts[0]
    return out


def gen_reads(n: int, seed: int, lmin: int = 1400, lmax: int = 1600, sites=SITES_16S):
    """C2-style full-length 16S reads: random ACGT backbone of length U[lmin,lmax] with every site
    planted in gene order.  Returns (arena u8[], offsets u64[n+1]).  Fully vectorised."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lmin, lmax + 1, size=n, dtype=np.int64)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:].view(np.int64))
    arena = random_acgt(rng, int(offsets[-1]))
    starts = offsets[:-1].astype(np.int64)
    for coord, site in sites:
        pos = starts + (lens * (coord / ECOLI_LEN)).astype(np.int64)
        pos = np.minimum(pos, starts + lens - len(site))
        mat = _resolve(rng, site, n)
        idx = pos[:, None] + np.arange(len(site), dtype=np.int64)[None, :]
        arena[idx] = mat
    return arena, offsets


def _edit(rng, s: bytearray, n_edits: int) -> bytearray:
    for _ in range(n_edits):
        kind = rng.integers(0, 3)
        p = int(rng.integers(0, max(1, len(s))))
        c = b"ACGT"[int(rng.integers(0, 4))]
        if kind == 0 and len(s):
            s[p] = c
        elif kind == 1:
            s.insert(p, c)
        elif len(s) > 1:
            del s[p]
    return s


def gen_reads_mutated(n: int, seed: int, max_edits: int = 3, revcomp_frac: float = 0.5, lower_frac: float = 0.0,
                      rna_frac: float = 0.0, lmin: int = 1400, lmax: int = 1600, sites=SITES_16S):
    """C3-style reads: every planted site gets 0..max_edits random sub/ins/del edits, a fraction of the
    reads is reverse-complemented, optionally soft-masked (lower case) or transcribed to RNA."""
    rng = np.random.default_rng(seed)
    recs = []
    for _ in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        seq = bytearray(random_acgt(rng, L).tobytes())
        for coord, site in sites:
            s = bytearray(_resolve(rng, site, 1)[0].tobytes())
            s = _edit(rng, s, int(rng.integers(0, max_edits + 1)))
            p = min(int(L * coord / ECOLI_LEN), max(0, L - len(s)))
            seq[p:p + len(s)] = s
        a = np.frombuffer(bytes(seq), dtype=np.uint8).copy()
        if rng.random() < revcomp_frac:
            a = revcomp_bytes(a)
        if rng.random() < rna_frac:
            a[a == ord("T")] = ord("U")
        if lower_frac > 0 and rng.random() < lower_frac:
            lo, hi = sorted(rng.integers(0, len(a) + 1, size=2))
            a[lo:hi] |= 0x20
        recs.append(a)
    return pack(recs)


def gen_contigs(n: int, seed: int, length: int = 5_000_000, copies=(3, 7), mask_frac=(0.3, 0.5), n_runs: int = 3,
                max_edits: int = 2):
    """C4-style soft-masked assemblies: random ACGT contigs with several (possibly reverse-strand, mutated)
    rRNA operon copies, lower-cased runs and a few N runs."""
    rng = np.random.default_rng(seed)
    recs = []
    for _ in range(n):
        a = random_acgt(rng, length)
        ncopy = int(rng.integers(copies[0], copies[1] + 1))
        for _c in range(ncopy):
            op, _ = gen_reads_mutated(1, int(rng.integers(0, 2**31)), max_edits=max_edits, revcomp_frac=0.3)
            if len(op) + 1 >= length:
                continue
            p = int(rng.integers(0, length - len(op)))
            a[p:p + len(op)] = op
        frac = rng.uniform(*mask_frac)
        run = max(1, length // 50)
        nmask = int(frac * length / run)
        for _m in range(nmask):
            p = int(rng.integers(0, max(1, length - run)))
            a[p:p + run] |= 0x20
        for _r in range(n_runs):
            ln = int(rng.integers(1, 200))
            p = int(rng.integers(0, max(1, length - ln)))
            a[p:p + ln] = ord("N")
        recs.append(a)
    return pack(recs)


def gen_primer_pairs(n_random: int, seed: int, lmin: int = 15, lmax: int = 40, builtins=None):
    """C5-style primer CSV content: the built-in pairs followed by n_random seeded IUPAC primer pairs."""
    rng = np.random.default_rng(seed)
    letters = "ACGT" * 6 + "MRWSYKVHDBN"
    pairs = [list(p) for p in (builtins or [])]
    for _ in range(n_random):
        pr = []
        for _s in range(2):
            L = int(rng.integers(lmin, lmax + 1))
            pr.append("".join(letters[int(i)] for i in rng.integers(0, len(letters), size=L)))
        pairs.append(pr)
    return pairs


def pack(records):
    """list of uint8 arrays / bytes -> (arena, offsets)"""
    arrs = [np.frombuffer(r, dtype=np.uint8) if isinstance(r, (bytes, bytearray)) else np.asarray(r, dtype=np.uint8)
            for r in records]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([len(a) for a in arrs])
    arena = np.concatenate(arrs) if arrs and int(offsets[-1]) else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(arena), offsets


def to_fasta(arena, offsets, ids=None, width: int = 0, crlf: bool = False) -> bytes:
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(len(offsets) - 1):
        s = arena[int(offsets[i]):int(offsets[i + 1])].tobytes()
        name = ids[i] if ids else f"rec{i}"
        out.append(b">" + name.encode() + nl)
        if width:
            for j in range(0, len(s), width):
                out.append(s[j:j + width] + nl)
        else:
            out.append(s + nl)
    return b"".join(out)


def write_fasta_file(path: str, arena: np.ndarray, offsets: np.ndarray, prefix: str = "r") -> int:
    """plain FASTA, one sequence line per record, ids <prefix><index>; returns the file size (bench.py's file_e2e input)"""
    mv = memoryview(np.ascontiguousarray(arena))
    off = [int(x) for x in offsets]
    pre = prefix.encode()
    with open(path, "wb", buffering=8 << 20) as f:
        for i in range(len(off) - 1):
            f.write(b">%s%d\n" % (pre, i))
            f.write(mv[off[i]:off[i + 1]])
            f.write(b"\n")
        return f.tell()
