"""Parity at the BASELINE.json sizes (test infrastructure; run under gpurun / torchrun).

For one workload (c2 | c3 | c4 | c5) every rank generates its shard exactly as bench.py does, runs the device path
through the C ABI (hx_run_batch on host buffers), and checks
  * a seeded sample of records against the CPU oracle (bit-exact), and
  * size-independent properties of the full table (positions inside the record, f_start < r_end, flags consistent);
rank 0 prints one JSON line with the per-rank sha256 of the full result tables (the "hit-table checksum" of
BASELINE.md row 5).   python tools/verify_at_scale.py --workload c5 --records 1250000 --sample 300
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402

import bench  # noqa: E402
import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--records", type=int, default=1_000_000)
    ap.add_argument("--sample", type=int, default=1000)
    ap.add_argument("--k", type=int, default=-1)
    ap.add_argument("--fast", type=int, default=1, help="library option fast_small_k (1 = default)")
    args = ap.parse_args()
    rank, world, local = bench.env_int("RANK", 0), bench.env_int("WORLD_SIZE", 1), bench.env_int("LOCAL_RANK", 0)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    arena, offsets, pairs, k, desc = bench.make_workload(args.workload, args.records, rank, args.k)
    n = len(offsets) - 1
    t0 = time.perf_counter()
    with hx.Context((local,)) as ctx:
        ctx.set_option("fast_small_k", args.fast)
        ctx.set_primers(pairs, k)
        res = ctx.run_batch(arena, offsets)
        groups = ctx.group_info()
    dt = time.perf_counter() - t0
    # properties of the full table
    lens = np.diff(offsets.astype(np.int64))
    valid = (res["flags"] & 4).astype(bool)
    assert ((res["flags"] & 3)[valid] == 3).all(), "a pair without both primers found"
    assert (res["r_end"][valid] < np.broadcast_to(lens[:, None], res.shape)[valid]).all(), "r_end outside the record"
    assert (res["f_start"][valid] <= res["r_end"][valid]).all(), "empty region"
    assert (res["combined"][valid] <= 2 * k).all(), "combined distance above 2k"
    # sample against the oracle
    rng = np.random.default_rng(1234 + rank)
    idx = np.sort(rng.choice(n, size=min(args.sample, n), replace=False))
    sub_a, sub_o = hx_synth.pack([arena[int(offsets[i]):int(offsets[i + 1])] for i in idx])
    ref = O.run_batch(sub_a, sub_o, pairs, k, nthreads=os.cpu_count() or 1)
    ok = ref.tobytes() == res[idx].tobytes()
    mine = {"rank": rank, "records": n, "bases": int(offsets[-1]), "regions": int(valid.sum()),
            "sample_records": int(len(idx)), "sample_equal_oracle": bool(ok),
            "table_sha256": hashlib.sha256(res.tobytes()).hexdigest(), "run_batch_s": round(dt, 3)}
    allr = [mine]
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    if rank == 0:
        print(json.dumps({"workload": desc, "k": k, "fast_small_k": args.fast, "pairs": len(pairs), "groups": groups, "world": world,
                          "all_samples_equal": all(r["sample_equal_oracle"] for r in allr), "ranks": allr}))
    if world > 1:
        dist.destroy_process_group()
    assert ok, "GPU result table differs from the oracle on the sample"


if __name__ == "__main__":
    main()
