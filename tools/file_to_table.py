"""FASTA file -> result table through the library's reader (hx_fasta_*) and batch calls, over reader thread counts and
chunk sizes: python tools/file_to_table.py [records] > gpurun_out/file_to_table.json   (needs a GPU)"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
arena, offsets = hx_synth.gen_reads(n, seed=2)
bases = int(offsets[-1])
work = bench.scratch_dir(int(bases * 1.2) + (1 << 30))
try:
    path = os.path.join(work, "in.fa")
    hx_synth.write_fasta_file(path, arena, offsets)
    pairs = hx.default_primers()
    res = []
    for chunk_mb in (16, 32, 64):
        variants = [("one_thread_ascii", 1, False)] + [(f"ahead{t}_ascii", t, False) for t in (2, 4, 8, 12)] + \
                   [(f"ahead{t}_packed", t, True) for t in (4, 8, 12)]
        r = bench.file_to_table_leg(hx, (0,), path, pairs, 0, n, bases, 2, chunk_mb << 20, variants)
        res.append({"chunk_mb": chunk_mb, **{v[0]: r[v[0]]["best"] for v in variants}, "tables_equal": r["tables_equal"]})
        print(json.dumps(res[-1]), flush=True)
finally:
    shutil.rmtree(work, ignore_errors=True)
