"""hx_run_batch host->device modes side by side on one GPU: ASCII slabs (0), packed on the way (1), both at once (2).
    python tools/mixed_ab.py [records] [reps]      (HYPEREX_HYB_DEBUG=1 prints the mixed path's scheduler counters)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import hx_synth  # noqa: E402
import hyperex_b200 as hx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
a, o = hx_synth.gen_reads(n, seed=2)
bases = int(o[-1])
h_arena = hx.pinned_empty(bases + 64)
h_arena[:bases] = a[:bases]
pairs = hx.default_primers()
with hx.Context((0,)) as ctx:
    ctx.set_primers(pairs, 0)
    h_out = hx.pinned_empty(n * len(pairs) * 12).view(hx.RESULT_DTYPE).reshape(n, len(pairs))
    ref = None
    cases = [(0, 0), (1, 0), (2, 0), (2, 12), (2, 8), (1, 8), (2, 4), (1, 4), (2, 0)]
    if os.environ.get("HX_AB_CASES"):  # "mode:threads,mode:threads"
        cases = [tuple(int(x) for x in c.split(":")) for c in os.environ["HX_AB_CASES"].split(",")]
    for mode, threads in cases:
        ctx.set_option("pack_h2d", mode)
        ctx.set_option("pack_threads", threads)
        for _ in range(3):
            ctx.run_batch(h_arena[:bases], o, out=h_out)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            ctx.run_batch(h_arena[:bases], o, out=h_out)
            ts.append(time.perf_counter() - t0)
        st = ctx.batch_stats()
        if ref is None:
            ref = h_out.tobytes()
        assert h_out.tobytes() == ref, (mode, threads)
        ts.sort()
        print(json.dumps({"pack_h2d": mode, "pack_threads": threads, "ms_median": ts[len(ts) // 2] * 1e3, "ms_min": ts[0] * 1e3,
                          "gbases_per_s_median": bases / ts[len(ts) // 2] / 1e9, "slabs": st["slabs"], "slabs_packed": st["slabs_packed"],
                          "h2d_bytes": st["h2d_bytes"], "linkq": os.environ.get("HYPEREX_MIXED_LINKQ", "3"),
                          "chunk_mb": os.environ.get("HYPEREX_MIXED_CHUNK_MB", "8")}), flush=True)
