"""Positions beyond 2^31 and the record-length limit (test infrastructure; run under gpurun, needs ~12 GB of host RAM).

  * one record of 2^31 + 150 MB of random ACGT with primer sites planted at its start, across the 2^31 boundary and at
    its end: the device path (hx_run_batch on host buffers, windows) must equal the CPU oracle bit for bit;
  * one record of 2^32 bytes: HX_ERR_ARG, never a truncated answer.
python tools/verify_big_record.py   -> one JSON line
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402

import hyperex_b200 as hx  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    L = (1 << 31) + 150_000_000
    rng = np.random.default_rng(31)
    arena = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L, dtype=np.uint8)]
    pairs = [["GTGCCAGCMGCCGCGGTAA", "GGACTACHVGGGTWTCTAAT"], ["CCTACGGGNGGCWGCAG", "GACTACHVGGGTATCTAATCC"]]  # v4, v3v4
    f = [b"GTGCCAGCAGCCGCGGTAA", b"CCTACGGGAGGCAGCAG"]
    r = [hx.to_reverse_complement("GGACTACAAGGGTATCTAAT", "dna").encode(),
         hx.to_reverse_complement("GACTACACGGGTATCTAATCC", "dna").encode()]

    def plant(pos, s):
        arena[pos:pos + len(s)] = np.frombuffer(s, dtype=np.uint8)

    # second pair early, first pair straddling 2^31, a forward-only site at the very end
    plant(1000, f[1]); plant(1400, r[1])
    plant((1 << 31) - 9, f[0]); plant((1 << 31) + 300, r[0])
    plant(L - len(f[1]), f[1])
    offsets = np.array([0, L], dtype=np.uint64)
    out = {"record_bytes": L}
    with hx.Context((0,)) as ctx:
        for k in (0, 1):
            ctx.set_primers(pairs, k)
            t0 = time.perf_counter()
            res = ctx.run_batch(arena, offsets)
            t1 = time.perf_counter()
            ref = O.run_batch(arena, offsets, pairs, k, nthreads=1)
            t2 = time.perf_counter()
            assert res.tobytes() == ref.tobytes(), (k, res, ref)
            out[f"k{k}"] = {"gpu_s": t1 - t0, "oracle_s": t2 - t1, "result": [[int(x["f_start"]), int(x["r_end"]),
                                                                              int(x["flags"]), int(x["combined"])]
                                                                             for x in res[0]]}
        assert res[0][0]["f_start"] == (1 << 31) - 9 and res[0][0]["r_end"] == (1 << 31) + 319
        del arena
        big = np.zeros(1 << 32, dtype=np.uint8)
        big[:] = ord("A")
        try:
            ctx.run_batch(big, np.array([0, 1 << 32], dtype=np.uint64))
            out["too_long"] = "accepted (BUG)"
        except hx.HyperexError as e:
            out["too_long"] = f"rejected: {e}"
    print(json.dumps(out))
    assert out["too_long"].startswith("rejected")


if __name__ == "__main__":
    main()
