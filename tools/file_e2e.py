"""The file-level leg of bench.py on its own (FASTA file in -> .fa + .gff out through hx_get_hypervar_regions):
   python tools/file_e2e.py [--records N] [--gpus G] [--steps S] [--k K]
Environment knobs of the library apply (HYPEREX_WRITE=pwrite, HYPEREX_THREADS, HYPEREX_CHUNK_BYTES, HYPEREX_TIMING=1)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import bench  # noqa: E402
import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=1_000_000)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--k", type=int, default=0)
ap.add_argument("--dir", default="")
a = ap.parse_args()
arena, offsets = hx_synth.gen_reads(a.records, seed=2)
out = bench.file_e2e_leg(hx, range(a.gpus), arena, offsets, hx.default_primers(), a.k, a.steps, a.dir)
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("HYPEREX_")}
print(json.dumps(out))
