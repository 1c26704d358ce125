"""One process driving several GPUs through hx_create(dev_ids, n): result table must equal the oracle's."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch  # noqa: E402

import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402
from oracle import oracle as O  # noqa: E402

nd = torch.cuda.device_count()
a, o = hx_synth.gen_reads_mutated(6000, seed=11, max_edits=2, revcomp_frac=0.3, lower_frac=0.2, rna_frac=0.1)
pairs = hx.default_primers()
ref = O.run_batch(a, o, pairs, 2, nthreads=8)
with hx.Context(tuple(range(nd))) as ctx:
    ctx.set_option("slab_bytes", 1 << 20)  # many slabs so every device gets work
    ctx.set_primers(pairs, 2)
    res = ctx.run_batch(a, o)
print("devices", nd, "equal", res.tobytes() == ref.tobytes())
assert res.tobytes() == ref.tobytes()
