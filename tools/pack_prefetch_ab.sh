for pf in ${PF_LIST:-0 -2048 -4096 -8192 -16384 0 -3072}; do
  HYPEREX_PACK_PF=$pf python - <<PY
import sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, hx_synth, hyperex_b200 as hx
n = 1000000
a, o = hx_synth.gen_reads(n, seed=2); bases = int(o[-1])
h = hx.pinned_empty(bases + 64); h[:bases] = a[:bases]
pk = hx.pinned_empty(bases // 2 + 64); fl = hx.pinned_empty(n)
res = {}
for th in (1, 16):
    hx.pack_records(h[:bases], o, th, pk, fl)
    t0 = time.perf_counter()
    for _ in range(3): hx.pack_records(h[:bases], o, th, pk, fl)
    res[th] = bases / ((time.perf_counter() - t0) / 3) / 1e9
pairs = hx.default_primers()
with hx.Context((0,)) as ctx:
    ctx.set_primers(pairs, 0)
    out = hx.pinned_empty(n * len(pairs) * 12).view(hx.RESULT_DTYPE).reshape(n, len(pairs))
    ctx.set_option("pack_h2d", 1); ctx.set_option("slab_bytes", 256 << 20)
    for _ in range(2): ctx.run_batch(h[:bases], o, out=out)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); ctx.run_batch(h[:bases], o, out=out); ts.append(time.perf_counter() - t0)
    ts.sort()
print("pf", os.environ["HYPEREX_PACK_PF"], "pack alone 1 thr %.1f 16 thr %.1f Gbases/s; run_batch median %.2f ms min %.2f" % (res[1], res[16], ts[3]*1e3, ts[0]*1e3))
PY
done
