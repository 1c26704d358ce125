"""A/B of hx_run_batch's host->device path on the GPU box: ASCII slabs (pack_h2d = 0) against slabs packed to 4 bits per
base on the way (pack_h2d = 1), over pack thread counts and slab sizes; plus the packer alone at several thread counts.
    python tools/pack_h2d_ab.py [records] > gpurun_out/pack_h2d_ab.json
Both paths must produce the same result table (asserted)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402

import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    a, o = hx_synth.gen_reads(n, seed=2)
    bases = int(o[-1])
    h_arena = hx.pinned_empty(bases + 64)
    h_arena[:bases] = a[:bases]
    pairs = hx.default_primers()
    out = {"records": n, "bases": bases, "k": k, "host_threads": os.cpu_count(), "pack_alone": [], "run_batch": []}
    pk = hx.pinned_empty(bases // 2 + 64)
    fl = hx.pinned_empty(n)
    for nt in (1, 0):
        os.environ["HYPEREX_PACK_NT"] = str(nt)  # non-temporal stores of the packed bytes (default for large batches) / ordinary
        for th in (1, 2, 4, 8, 12, 16, 24, 32):
            if th > (os.cpu_count() or 1):
                break
            hx.pack_records(h_arena[:bases], o, th, pk, fl)
            t0 = time.perf_counter()
            for _ in range(3):
                hx.pack_records(h_arena[:bases], o, th, pk, fl)
            dt = (time.perf_counter() - t0) / 3
            out["pack_alone"].append({"nt_stores": nt, "threads": th, "ms": dt * 1e3, "gbases_per_s": bases / dt / 1e9})
    del os.environ["HYPEREX_PACK_NT"]
    # do the packer (CPU, host memory) and a host->device copy (DMA out of host memory) slow each other down?
    import torch
    t_pk = torch.from_numpy(pk[:bases // 2])
    d_pk = torch.empty(bases // 2, dtype=torch.uint8, device="cuda:0")
    side = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        d_pk.copy_(t_pk, non_blocking=True)
        e0.record()
        for _ in range(4):
            d_pk.copy_(t_pk, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    out["contention"] = {"h2d_alone_gbs": 4 * (bases // 2) / (e0.elapsed_time(e1) * 1e-3) / 1e9}
    with torch.cuda.stream(side):
        e0.record()
        for _ in range(8):
            d_pk.copy_(t_pk, non_blocking=True)
        e1.record()
    time.sleep(0.002)
    t0 = time.perf_counter()
    for _ in range(4):
        hx.pack_records(h_arena[:bases], o, 0, pk, fl)
    dt = (time.perf_counter() - t0) / 4
    torch.cuda.synchronize()
    out["contention"].update({"pack_while_h2d_gbases_per_s": bases / dt / 1e9, "pack_while_h2d_ms": dt * 1e3,
                              "h2d_while_pack_gbs": 8 * (bases // 2) / (e0.elapsed_time(e1) * 1e-3) / 1e9,
                              "note": "8 copies of the packed arena (about 110 ms) on a side stream, 4 pack calls (all threads) meanwhile"})
    with hx.Context((0,)) as ctx:
        ctx.set_primers(pairs, k)
        h_out = hx.pinned_empty(n * len(pairs) * 12).view(hx.RESULT_DTYPE).reshape(n, len(pairs))
        ref = None
        for mode, threads, slab_mb, piece_kb in [(0, 0, 128, 4096), (1, 0, 128, 4096), (1, 0, 128, 1024), (1, 0, 128, 2048), (1, 0, 128, 8192),
                                                 (1, 0, 128, 16384), (1, 0, 128, 65536), (1, 0, 256, 4096), (1, 0, 64, 4096), (1, 12, 128, 4096),
                                                 (1, 8, 128, 4096), (1, 4, 128, 4096), (-1, 0, 128, 4096)]:
            ctx.set_option("pack_piece_bytes", piece_kb << 10)
            ctx.set_option("pack_h2d", mode)
            ctx.set_option("pack_threads", threads)
            ctx.set_option("slab_bytes", slab_mb << 20)
            for _ in range(2):
                ctx.run_batch(h_arena[:bases], o, out=h_out)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                ctx.run_batch(h_arena[:bases], o, out=h_out)
                ts.append(time.perf_counter() - t0)
            if ref is None:
                ref = h_out.tobytes()
            assert h_out.tobytes() == ref, "pack_h2d changed the result table"
            ts.sort()
            out["run_batch"].append({"pack_h2d": mode, "pack_threads": threads, "slab_mb": slab_mb, "piece_kb": piece_kb, "ms_median": ts[2] * 1e3,
                                     "ms_min": ts[0] * 1e3, "gbases_per_s_median": bases / ts[2] / 1e9})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
