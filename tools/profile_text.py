"""Drives the on-device FASTA / GFF3 assembly (hx_run_batch_text) on one CLI-sized batch, for ncu:
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/profile_text.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402

n = 44_000
a, o = hx_synth.gen_reads(n, seed=3)
ids = [f"r{i}" for i in range(n)]
pairs = hx.default_primers()
fa_tails, gff_tails = [], []
for f, r in pairs:
    lab = hx.primers_to_region([f, r])
    fa_tails.append((f" region={lab}" if lab else "") + f" forward={f} reverse={r}\n")
    gff_tails.append("\t.\t.\t.\tNote=" + (f"Hypervariable region {lab}" if lab else f"forward={f} reverse={r}") + "\n")
with hx.Context((0,)) as ctx:
    ctx.set_primers(pairs, 0)
    for _ in range(2):
        res, fa, gff = ctx.run_batch_text(a, o, ids, fa_tails, gff_tails, stream=True)
print("bases", int(o[-1]), "fasta_bytes", len(fa), "gff_bytes", len(gff))
