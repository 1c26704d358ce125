"""Vectors made by the REFERENCE itself (rust-bio 1.6.0 and the unmodified hyperex v0.2.0 binary), when present.

They cannot be generated in the build image (no Rust toolchain, no network): oracle/ref_recipe/run.sh produces them on
any machine with cargo and writes tests/golden/ref/.  Until someone has run it, these tests skip and parity stays
"unpinned against the real reference" (oracle/hx_oracle.h); once the files exist, the oracle — and through
it every GPU parity test — is pinned to the reference's own answers."""
import json
import os

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_case_list_is_committed_and_answerable_by_the_oracle():
    rows = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(GOLDEN, "ref_cases.tsv"))]
    assert len(rows) > 300 and all(len(r) == 3 for r in rows)
    # the five rust-bio documentation vectors lead the list; first one with hyperex's ambiguity table
    assert O.find_all_end(rows[0][0], rows[0][1], int(rows[0][2])) == [(11, 2), (12, 2)]
    for p, t, k in rows[:50]:
        assert O.find_all_end(p, t, int(k)) == O.dp_find_all_end(p, t, int(k))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "find_all_lazy.jsonl")),
                    reason="reference-generated vectors absent: run oracle/ref_recipe/run.sh where cargo exists")
def test_oracle_equals_rust_bio_find_all_lazy():
    n = 0
    for line in open(os.path.join(REF, "find_all_lazy.jsonl")):
        c = json.loads(line)
        assert O.find_all_end(c["pattern"], c["text"], c["k"]) == [tuple(h) for h in c["hits"]], c["pattern"]
        n += 1
    assert n > 300


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "config1_k0.fa")),
                    reason="reference-generated outputs absent: run oracle/ref_recipe/run.sh where cargo exists")
def test_oracle_bytes_equal_the_reference_binary():
    for name in ("config1_k0", "config1_k2"):
        for ext in ("fa", "gff"):
            assert open(os.path.join(REF, f"{name}.{ext}"), "rb").read() == open(os.path.join(GOLDEN, f"{name}.{ext}"), "rb").read()
    for inp in sorted(os.listdir(os.path.join(GOLDEN, "ref_inputs"))):
        buf = open(os.path.join(GOLDEN, "ref_inputs", inp), "rb").read()
        base = inp[:-3]
        for tag, pairs, k in ((f"{base}_k0", O.default_pairs(), 0), (f"{base}_k2", O.default_pairs(), 2),
                              (f"{base}_v3v4_v4v5_k3", [O.region_to_primer("v3v4"), O.region_to_primer("v4v5")], 3)):
            fa, gff, _ = O.get_hypervar_regions(buf, pairs, k)
            assert open(os.path.join(REF, tag + ".fa"), "rb").read() == fa, tag
            assert open(os.path.join(REF, tag + ".gff"), "rb").read() == gff, tag
