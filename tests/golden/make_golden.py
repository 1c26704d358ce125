"""Regenerates the committed golden vectors.

The reference cannot be built here (Rust; rust-bio un-vendored), so these files are
produced by the CPU oracle (oracle/hx_oracle.c), which is itself pinned by the rust-bio doc vectors
and the independent DP in tests/test_oracle.py.  The config-1 outputs reproduce the hashes that
SURVEY.md §8c derived independently.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]

import numpy as np  # noqa: E402

import hx_synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    buf = open(os.path.join(HERE, "test.fa"), "rb").read()
    meta = {}
    for k in (0, 2):
        fa, gff, res = O.get_hypervar_regions(buf, O.default_pairs(), k)
        open(os.path.join(HERE, f"config1_k{k}.fa"), "wb").write(fa)
        open(os.path.join(HERE, f"config1_k{k}.gff"), "wb").write(gff)
        meta[f"config1_k{k}"] = {"fa_sha256": hashlib.sha256(fa).hexdigest(),
                                 "gff_sha256": hashlib.sha256(gff).hexdigest(), "fa_bytes": len(fa),
                                 "gff_bytes": len(gff)}
    # hits of every built-in pattern on test.fa (k = 0..3), the per-pattern table of SURVEY §8c
    arena, offsets, ids = O.parse_fasta(buf)
    text = arena.tobytes()
    hits = {}
    for region, (f, r) in zip(O.supported_regions(), O.default_pairs()):
        for name, pat in ((f"F:{f}", f), (f"R:{r}", O.to_reverse_complement(r, O.DNA).decode())):
            for k in range(4):
                hits[f"{name}|k{k}"] = O.find_all_end(pat, text, k)
    meta["test_fa_hits"] = hits
    # small seeded batches with their oracle result tables (C2/C3-like, plus edge cases)
    cases = {}
    a, o = hx_synth.gen_reads(64, seed=2)
    cases["c2_small"] = (a, o, O.default_pairs(), 0)
    a, o = hx_synth.gen_reads_mutated(64, seed=3, max_edits=3, revcomp_frac=0.5)
    cases["c3_small"] = (a, o, [O.region_to_primer("v3v4"), O.region_to_primer("v4v5")], 2)
    a, o = hx_synth.gen_reads_mutated(48, seed=4, max_edits=2, revcomp_frac=0.2, lower_frac=0.5, rna_frac=0.3)
    cases["mixed_small"] = (a, o, O.default_pairs(), 3)
    for name, (a, o, pairs, k) in cases.items():
        res = O.run_batch(a, o, pairs, k)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), arena=a, offsets=o, res=res,
                            pairs=np.array(pairs), k=k)
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps({k: v for k, v in meta.items() if k != "test_fa_hits"}, indent=1))


if __name__ == "__main__":
    main()
