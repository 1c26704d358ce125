#!/usr/bin/env python
"""bench.py — HyperEx primer-search-and-extract hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (classify -> Myers scan of every forward primer and every
reverse-complemented reverse primer -> pairing) over one batch of synthetic records.  Workload =
BASELINE.json configs[1]: 1 M synthetic full-length 16S reads (~1.5 kb, every built-in primer site
planted), all 10 built-in regions, -m 0, per GPU.  With N > 1 (torchrun, one rank per GPU) every rank
scans its own shard of N x 1 M records; no collective is on the data path, timing is the
max over ranks.

  value        Gbases/s with the arena resident in HBM (CUDA events on the launching stream)
  e2e          Gbases/s through hx_run_batch with HOST buffers: pinned arena -> H2D -> kernels -> D2H
               result table, every step (library defaults: the small-k fast path is on, as shipped; how the text
               crosses the link — ASCII slabs, or ASCII and packed slabs at once — is what the library measured,
               e2e.mode; ascii_slabs / packed_slabs / mixed_slabs are the forced modes next to it)
  file_e2e     Gbases/s of the drop-in seam itself, hx_get_hypervar_regions: FASTA file in, <prefix>.fa and
               <prefix>.gff out (utils.rs:231-414 / main.rs:38), with the parse / device / write phases;
               file_e2e.file_to_table: the same file -> host result table through hx_fasta_next* + hx_run_batch*
  roofline     the Myers kernel against the measured integer-ALU peak (hx_int_peak) and HBM peak
  cpu_baseline the CPU oracle (C port of the reference path) on a bounded sample, one thread like the
               single-threaded reference
  --impl reference   the same port on all host threads, as the reference arm
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

CELLS_PER_BASE_DEFAULT = 393  # sum over the 10 built-in pairs of len(fwd)+len(rev)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=1_000_000, help="records per GPU")
    ap.add_argument("--k", type=int, default=-1, help="-m value (default: the workload's own)")
    ap.add_argument("--long-filter", type=int, default=1, choices=[0, 1],
                    help="1 (library default): primers > 32 nt are scanned as 32-bit suffix automata + verification; "
                         "0: 64-bit automata (only matters for c5)")
    ap.add_argument("--group-max", type=int, default=0, help="experiment: cap on patterns per group (0 = automatic)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 is the BASELINE.json metric config; c3/c4/c5 are informational")
    ap.add_argument("--cpu-sample", type=int, default=60_000, help="records of the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=0, help="records per step of the reference arm (0: auto)")
    ap.add_argument("--slab-mb", type=int, default=0, help="host->device pipeline granule of the e2e leg (0: library default)")
    ap.add_argument("--pack-h2d", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="hx_run_batch of the e2e leg: -1 library default (automatic), 0 ASCII slabs, 1 slabs packed to 4 bits per "
                         "base on their way to the device, 2 ASCII slabs from the front and packed slabs from the back at once")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ab", action="store_true", help="skip the TMA / LDG text-path A/B")
    ap.add_argument("--no-file-e2e", action="store_true", help="skip the file-level leg (FASTA file in, two files out)")
    ap.add_argument("--file-steps", type=int, default=3, help="timed calls of the file-level leg")
    ap.add_argument("--file-dir", default="", help="where the file-level leg keeps its files (default: /dev/shm if it has room)")
    return ap.parse_args()


def pairs_and_cells():
    from oracle import oracle as O  # only for the primer table of the CPU legs; the product has its own
    pairs = O.default_pairs()
    return pairs, sum(len(f) + len(r) for f, r in pairs)


def make_workload(name, n, rank, k_arg):
    """(arena, offsets, pairs, k, description) for one rank; shapes follow BASELINE.json configs[1..4].
    The primer tables come from the product's own host helpers (the oracle is only used by the CPU legs)."""
    import hx_synth
    import hyperex_b200 as O  # default_primers / region_to_primer / to_reverse_complement of the product
    if name == "c2":
        a, o = hx_synth.gen_reads(n, seed=2 + 1000 * rank)
        return a, o, O.default_primers(), 0 if k_arg < 0 else k_arg, (
            "c2: 1M synthetic full-length 16S reads ~1.5 kb per GPU, 10 built-in regions (20 searches/base, "
            "393 cells/base), -m 0")
    if name == "c3":
        base_n = min(n, 20000)  # the mutated generator is a Python loop: make 20k reads and repeat them
        a, o = hx_synth.gen_reads_mutated(base_n, seed=3 + 1000 * rank, max_edits=3, revcomp_frac=0.5)
        reps = max(1, n // base_n)
        lens = np.tile(np.diff(o.astype(np.int64)), reps)
        a = np.tile(a, reps)
        o = np.zeros(len(lens) + 1, np.uint64)
        o[1:] = np.cumsum(lens)
        return a, o, [O.region_to_primer("v3v4"), O.region_to_primer("v4v5")], 2 if k_arg < 0 else k_arg, (
            f"c3: {len(lens)} reads with IUPAC-degenerate primers v3v4+v4v5, sites with 0-3 edits, half reverse-"
            "complemented, -m 2")
    if name == "c4":
        nc = max(1, n // 15625)  # n = 1M -> 64 contigs of 5 Mb
        a, o = hx_synth.gen_contigs(nc, seed=4 + 1000 * rank, length=5_000_000)
        return a, o, O.default_primers(), 3 if k_arg < 0 else k_arg, (
            f"c4: {nc} soft-masked 5 Mb contigs with multi-copy rRNA operons, 10 built-in regions, -m 3, windows")
    a, o = hx_synth.gen_reads(n, seed=5 + 1000 * rank)
    pairs = hx_synth.gen_primer_pairs(54, seed=5, builtins=O.default_primers())
    return a, o, pairs, 2 if k_arg < 0 else k_arg, (
        f"c5: {n} SILVA-scale 16S records per GPU + 64-pair primer CSV (some primers > 32 nt), -m 2")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("HX_BENCH_SMI_MS", "100")], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(bases):
    """dram__bytes_read.sum + dram__bytes_write.sum of hx_myers_pair_kernel from the committed ncu capture of
    this same command (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py from the ncu export); only
    reported when the capture was taken on the same number of bases per launch."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            if int(t["bases_per_launch"]) == int(bases):
                return {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "dram_bytes_read": t["dram_bytes_read"],
                        "dram_bytes_write": t["dram_bytes_write"], "algorithmic_bytes": int(bases),
                        "source": t.get("source", "") + f" [profiles/{name}]"}
        except Exception:
            continue
    return None


def cpu_port_run(arena, offsets, n, pairs, k, nthreads, rebuild=True):
    """times the oracle (C port of utils.rs:270-351 + rust-bio Myers) on the first n records"""
    from oracle import oracle as O
    sub_o = np.ascontiguousarray(offsets[:n + 1])
    sub_a = arena[:int(sub_o[-1])]
    t0 = time.perf_counter()
    O.run_batch(sub_a, sub_o, pairs, k, nthreads=nthreads, rebuild_per_record=rebuild)
    dt = time.perf_counter() - t0
    return int(sub_o[-1] - sub_o[0]), dt


def scratch_dir(need_bytes, override=""):
    """directory for the file-level leg: /dev/shm when it has room (the leg measures the pipeline, not a disk)"""
    import shutil
    import tempfile
    for d in ([override] if override else []) + ["/dev/shm", tempfile.gettempdir(), os.path.join(ROOT, "gpurun_out")]:
        try:
            os.makedirs(d, exist_ok=True)
            if shutil.disk_usage(d).free > need_bytes:
                return tempfile.mkdtemp(prefix="hx_file_e2e_", dir=d)
        except OSError:
            continue
    return None


def sha256_file(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(16 << 20)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def file_e2e_leg(hx, devices, arena, offsets, pairs, k, steps, work_override=""):
    """FASTA file in -> <prefix>.fa + <prefix>.gff out through hx_get_hypervar_regions on a context over `devices`
    (what `bin/hyperex --gpus N` runs).  No log callback (the reference's per-record warn! lines are a side channel,
    SURVEY.md §2 row 3).  One untimed call first: it allocates the pinned batches the context then reuses."""
    import shutil
    import hx_synth
    n, bases = len(offsets) - 1, int(offsets[-1])
    work = scratch_dir(int(bases * 7.5) + (2 << 30), work_override)
    if work is None:
        return {"unavailable": "no scratch directory with room for the input and output files"}
    try:
        path = os.path.join(work, "in.fa")
        in_bytes = hx_synth.write_fasta_file(path, arena, offsets)
        prefix = os.path.join(work, "out")
        times, stats = [], None
        with hx.Context(tuple(devices)) as fctx:
            for it in range(steps + 1):
                for ext in (".fa", ".gff"):
                    if os.path.exists(prefix + ext):
                        os.remove(prefix + ext)
                t0 = time.perf_counter()
                fctx.get_hypervar_regions(path, pairs, prefix, k)
                dt = time.perf_counter() - t0
                if it:
                    times.append(dt)
                    if stats is None or dt <= min(times):
                        stats = fctx.file_stats()
        fa_bytes, gff_bytes = os.path.getsize(prefix + ".fa"), os.path.getsize(prefix + ".gff")
        mean = sum(times) / len(times)
        out = {"what": "hx_get_hypervar_regions: plain FASTA file -> .fa + .gff files, no log callback, host threads + "
                       f"{len(devices)} GPU(s) in one context", "value": bases / mean / 1e9, "unit": "Gbases/s",
               "best": bases / min(times) / 1e9, "seconds": times, "records": n, "bases": bases,
               "input_bytes": in_bytes, "fasta_bytes": fa_bytes, "gff_bytes": gff_bytes, "dir": os.path.dirname(work),
               "sha256": {"fa": sha256_file(prefix + ".fa"), "gff": sha256_file(prefix + ".gff")},
               "phases_of_best_call": {kk: stats[kk] for kk in ("wall_s", "setup_s", "chunk_s", "parse_s", "device_s", "write_s",
                                                              "close_s", "parsers", "lanes", "writers", "batches")},
               "phases_note": "stage seconds are summed over the threads of the stage (pipeline: they overlap)",
               "host_cores": os.cpu_count()}
        try:
            out["file_to_table"] = file_to_table_leg(hx, devices, path, pairs, k, n, bases, min(steps, 2))
        except Exception as e:  # pragma: no cover - reported, never hidden
            out["file_to_table"] = {"error": repr(e)}
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def file_to_table_leg(hx, devices, path, pairs, k, n, bases, steps, chunk=32 << 20, variants=None):
    """FASTA file -> result table on the host through the library's own reader and batch calls (what a caller that wants
    coordinates, not text, runs): hx_fasta_next + hx_run_batch on the calling thread (round 1's only way), the same with
    the chunks parsed ahead by background threads (hx_fasta_configure), and hx_fasta_next_packed + hx_run_batch_packed
    with the chunks parsed AND packed ahead (the producer emits the packed arena).  One table per variant, hashed."""
    import hashlib
    from hyperex_b200.api import RESULT_DTYPE
    raw = hx.pinned_empty(n * len(pairs) * RESULT_DTYPE.itemsize)
    table = raw.view(RESULT_DTYPE).reshape(n, len(pairs))
    variants = variants or (("one_thread_ascii", 1, False), ("read_ahead_ascii", 0, False), ("read_ahead_packed", 0, True))
    out = {"what": f"FASTA file -> host result table: library reader + hx_run_batch[_packed], {chunk >> 20} MiB chunks; the reader is opened and "
                   "closed inside the timed region, its pinned slots come from the process-wide cache the untimed first "
                   "pass filled (pinning costs 1.5 ms per MB on these hosts)", "unit": "Gbases/s", "host_cores": os.cpu_count()}
    try:
        with hx.Context(tuple(devices)) as tctx:
            tctx.set_primers(pairs, k)
            for name, threads, packed in variants:
                times, moved = [], None
                for it in range(steps + 1):
                    raw[:] = 0
                    t0 = time.perf_counter()
                    r0 = 0
                    if packed:
                        for pk, fl, off, _ids, _ar in hx.read_fasta_packed(path, chunk, threads=threads, pack_ahead=True,
                                                                           copy=False, want_ids=False):
                            nb = len(off) - 1
                            tctx.run_batch_packed(pk, fl, off, table[r0:r0 + nb])
                            r0 += nb
                    else:
                        for ar, off, _ids in hx.read_fasta(path, chunk, threads=threads, copy=False, want_ids=False):
                            nb = len(off) - 1
                            tctx.run_batch(ar, off, table[r0:r0 + nb])
                            r0 += nb
                    dt = time.perf_counter() - t0
                    if it:
                        times.append(dt)
                    moved = tctx.batch_stats()
                if r0 != n:
                    raise RuntimeError(f"{name}: {r0} records read, {n} written")
                mean = sum(times) / len(times)
                out[name] = {"value": bases / mean / 1e9, "best": bases / min(times) / 1e9, "seconds": times,
                             "reader_threads": threads if threads else max(1, min(8, (os.cpu_count() or 2) - 1)),
                             "text_crossed_packed": moved["packed_text"],
                             "table_sha256": hashlib.sha256(table.tobytes()).hexdigest()}
            out["tables_equal"] = len({out[v[0]]["table_sha256"] for v in variants}) == 1
    finally:
        hx.pinned_free(raw)
    return out


def run_reference(args, rank, world):
    """reference arm: the CPU port of the reference path on all host threads (rank 0 only)"""
    if rank != 0:
        return
    import hx_synth
    pairs, cells = pairs_and_cells()
    if args.k < 0:
        args.k = 0  # configs[1]: -m 0
    cores = os.cpu_count() or 1
    # The GPU arm's own record count per step when the host gets through it in time (probe: 20 000 records; budget
    # ~4 minutes for the whole run), else a bounded sample of the same generator and seed.
    full = args.records
    arena, offsets = hx_synth.gen_reads(full if not args.ref_sample else args.ref_sample, seed=2)
    probe_n = min(len(offsets) - 1, 20_000)
    cpu_port_run(arena, offsets, min(probe_n, 500 * cores), pairs, args.k, cores)  # warm-up (threads, page faults)
    pb, pdt = cpu_port_run(arena, offsets, probe_n, pairs, args.k, cores)
    n = len(offsets) - 1
    if not args.ref_sample:
        est_full = pdt * n / probe_n
        if est_full * (args.steps + 0.2) > 240.0:
            n = int(max(2000, min(n, probe_n * 240.0 / (args.steps + 0.2) / pdt)))
    same_config = n == args.records
    times, bases = [], 0
    for _ in range(args.steps):
        bases, dt = cpu_port_run(arena, offsets, n, pairs, args.k, cores)
        times.append(dt)
    dt = sum(times) / len(times)
    gb = bases / dt / 1e9
    line = {"impl": "reference", "metric": "Gbases/s", "value": gb, "unit": "Gbases/s", "gcups": gb * cells,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "c2: synthetic full-length 16S reads ~1.5 kb, 10 built-in regions, -m 0 "
                                   + (f"({n} records per step, the GPU arm's own count)" if same_config else
                                      f"(bounded sample of {n} records per step)"), "records": n, "pairs": len(pairs),
                       "k": args.k, "same_records_as_gpu_arm": same_config},
            "cpu_baseline": {"value": gb, "unit": "Gbases/s", "cores": cores, "kind": "port",
                             "sample": f"{n} records ({bases} bases) per step, {cores} threads; C port of the "
                                       "reference path (Rust toolchain absent: SURVEY.md §8c), Peq rebuilt per "
                                       "record x pair like utils.rs:301-302"},
            "e2e": {"value": gb, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_file_e2e:
        line["file_e2e"] = reference_file_leg(arena, offsets, min(n, 60_000), pairs, args.k, cores)
    print(json.dumps(line), flush=True)


def reference_file_leg(arena, offsets, n, pairs, k, cores):
    """the reference's seam on the host: FASTA file in -> FASTA + GFF3 files out with the C port (parse, search on all
    threads, format, write), on a bounded sample"""
    import shutil
    import hx_synth
    from oracle import oracle as O
    sub_o = np.ascontiguousarray(offsets[:n + 1])
    bases = int(sub_o[-1])
    work = scratch_dir(bases * 8 + (1 << 30))
    if work is None:
        return {"unavailable": "no scratch directory"}
    try:
        path = os.path.join(work, "in.fa")
        hx_synth.write_fasta_file(path, arena[:bases], sub_o)
        t0 = time.perf_counter()
        buf = open(path, "rb").read()
        a, o, ids = O.parse_fasta(buf)
        res = O.run_batch(a, o, pairs, k, nthreads=cores, rebuild_per_record=True)
        fa, gff = O.format_outputs(a, o, ids, pairs, res)
        with open(os.path.join(work, "out.fa"), "wb") as f:
            f.write(fa)
        with open(os.path.join(work, "out.gff"), "wb") as f:
            f.write(gff)
        dt = time.perf_counter() - t0
        return {"what": "C port: FASTA file -> parse -> search (all threads) -> format -> .fa + .gff files", "value": bases / dt / 1e9,
                "unit": "Gbases/s", "records": n, "bases": bases, "seconds": dt, "cores": cores, "fasta_bytes": len(fa),
                "gff_bytes": len(gff)}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    args = parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import hx_synth
    import hyperex_b200 as hx

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_cpus = hx.bind_to_gpu_numa_node(local) if world > 1 else None  # pinned arenas local to the GPU's PCIe root
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    # every rank owns a different shard of the N x records job (weak scaling); seed = config index + rank
    arena, offsets, pairs, k, workload_desc = make_workload(args.workload, args.records, rank, args.k)
    args.k = k
    cells = sum(len(f) + len(r) for f, r in pairs)
    n = len(offsets) - 1
    bases = int(offsets[-1])
    npairs = len(pairs)

    ctx = hx.Context((local,))
    # The graded kernel is the Myers automaton (north_star; SURVEY §8f-4: fast paths are "never counted as Myers
    # GCUPS"), so `value`, `e2e` and `roofline` are measured with the small-k fast path OFF; the library default
    # (fast path ON for -m 0..2) is timed separately below and reported as `fast_path`.
    ctx.set_option("fast_small_k", 0)
    ctx.set_option("long_filter", args.long_filter)
    ctx.set_option("group_np_max", args.group_max)
    if args.slab_mb > 0:
        ctx.set_option("slab_bytes", args.slab_mb << 20)
    ctx.set_primers(pairs, args.k)

    # ---- device-resident leg ("value") -------------------------------------------------------
    pad = (-bases) % 16 + 16
    d_arena = torch.empty(bases + pad, dtype=torch.uint8, device="cuda")
    d_arena[:bases].copy_(torch.from_numpy(arena))
    d_arena[bases:].zero_()
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    d_out = torch.empty(n * npairs * 12, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def dev_step():
        ctx.run_batch_device(d_arena.data_ptr(), d_off.data_ptr(), n, bases, d_out.data_ptr(),
                             stream=stream.cuda_stream, dev_slot=0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        dev_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.set_option("time_kernels", 1)
    ctx.kernel_time(reset=True)
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        dev_step()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    myers_ms, myers_launches = ctx.kernel_time(reset=True)
    ctx.set_option("time_kernels", 0)

    # ---- A/B of the text path (not part of the timed region): TMA bulk-copy staging vs the default ----
    ab_ms = None
    if rank == 0 and not args.no_ab:
        ctx.set_option("text_tma", 1)
        dev_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(2):
            dev_step()
        e1.record(stream)
        torch.cuda.synchronize()
        ab_ms = e0.elapsed_time(e1) / 2
        ctx.set_option("text_tma", 0)

    # ---- the library default for small k: Wu-Manber / Shift-Or fast path (not part of the timed region) ----
    fast_ms, fast_groups = None, (0, 0, 0)
    if args.k <= 2 and not args.no_ab:
        ctx.set_option("fast_small_k", 1)
        ctx.set_primers(pairs, args.k)  # the grouping depends on it (-m 2: groups of <= 8 automata)
        fast_groups = ctx.group_info()
        for _ in range(2):
            dev_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            dev_step()
        e1.record(stream)
        barrier()
        fast_ms = e0.elapsed_time(e1) / args.steps
        fast_res = d_out.cpu().numpy().tobytes()
        ctx.set_option("fast_small_k", 0)
        ctx.set_primers(pairs, args.k)
        dev_step()
        torch.cuda.synchronize()
        assert fast_res == d_out.cpu().numpy().tobytes(), "fast path and Myers path disagree"

    # ---- end-to-end leg ("e2e"): host arena in, host result table out, every step ----------------
    # through the public entry point as shipped: library defaults, i.e. the small-k fast path is ON for -m 0..2
    # (PCIe-bound either way; the result table is asserted equal to the Myers kernel's below)
    ctx.set_option("fast_small_k", 1)
    ctx.set_primers(pairs, args.k)
    h_arena = hx.pinned_empty(bases + 64)
    h_arena[:bases] = arena
    h_out_raw = hx.pinned_empty(n * npairs * 12)
    h_out = h_out_raw.view(hx.RESULT_DTYPE).reshape(n, npairs)
    # a rank shares the host's threads with the other ranks of the job: the packer of the packed-on-the-way path gets its share
    pack_threads = max(1, (os.cpu_count() or 1) // world)
    ctx.set_option("pack_threads", pack_threads)

    def e2e_leg(pack_h2d):
        """args.steps timed hx_run_batch calls, host arena -> host result table; returns (seconds, hx_batch_stats, sha-able bytes)"""
        ctx.set_option("pack_h2d", pack_h2d)
        # automatic: the library spends its first six large calls measuring both modes (all ranks at once, as in the
        # timed region); they are warm-up here, plus one call in the mode it settled on
        for _ in range(7 if pack_h2d < 0 else max(min(args.warmup, 3), 2)):
            ctx.run_batch(h_arena[:bases], offsets, out=h_out)
        barrier()
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            ctx.run_batch(h_arena[:bases], offsets, out=h_out)
        torch.cuda.synchronize()
        return time.perf_counter() - t0_, ctx.batch_stats(), h_out.tobytes()

    # `e2e` = the library default (pack_h2d automatic: the library measures ASCII slabs against the mixed path — ASCII slabs and
    # slabs packed to 4 bits per base on the way in one call — and keeps the faster); the forced modes are timed next to it
    e2e_s, e2e_stats, e2e_bytes = e2e_leg(args.pack_h2d)
    e2e_ascii_s, ascii_stats, ascii_bytes = e2e_leg(0)
    e2e_otw_s, otw_stats, otw_bytes = e2e_leg(1)
    e2e_hyb_s, hyb_stats, hyb_bytes = e2e_leg(2)
    assert e2e_bytes == ascii_bytes == otw_bytes == hyb_bytes, "pack_h2d changed the result table"
    ctx.set_option("pack_h2d", args.pack_h2d)
    ctx.run_batch(h_arena[:bases], offsets, out=h_out)
    clocks = sampler.stop()
    # context for e2e: the same arena as one plain pinned -> device copy (what PCIe alone costs), outside the timed region
    t_arena = torch.from_numpy(h_arena[:bases])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_arena[:bases].copy_(t_arena, non_blocking=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        d_arena[:bases].copy_(t_arena, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_only_ms = ev0.elapsed_time(ev1) / 3
    # ... and with every rank copying AT THE SAME TIME (N > 1): what the host can feed to N GPUs at once, i.e. the
    # ceiling of `e2e` on this box
    h2d_conc_ms = h2d_only_ms
    if world > 1:
        barrier()
        t0c = time.perf_counter()
        for _ in range(3):
            d_arena[:bases].copy_(t_arena, non_blocking=True)
        torch.cuda.synchronize()
        h2d_conc_ms = (time.perf_counter() - t0c) * 1e3 / 3

    # ---- packed-arena leg ("e2e_packed"): 4-bit text (north_star "2-bit/ASCII arena"), half the bytes over PCIe ---------
    # The packer (hx_pack_records, host, all threads) is timed on its own, and the leg is reported both without it (the
    # arena arrives packed, e.g. from a parser that packs while it touches the bytes anyway) and with it inside the timed
    # region every step.
    pk = hx.pinned_empty(bases // 2 + 64)
    fl = hx.pinned_empty(max(n, 1))
    h_out2_raw = hx.pinned_empty(n * npairs * 12)
    h_out2 = h_out2_raw.view(hx.RESULT_DTYPE).reshape(n, npairs)
    hx.pack_records(h_arena[:bases], offsets, 0, pk, fl)
    t0 = time.perf_counter()
    for _ in range(3):
        hx.pack_records(h_arena[:bases], offsets, 0, pk, fl)
    pack_s = (time.perf_counter() - t0) / 3
    for _ in range(min(args.warmup, 2)):
        ctx.run_batch_packed(pk, fl, offsets, out=h_out2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.run_batch_packed(pk, fl, offsets, out=h_out2)
    torch.cuda.synchronize()
    e2e_packed_s = time.perf_counter() - t0
    assert h_out2.tobytes() == h_out.tobytes(), "packed-arena and ASCII legs disagree"
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hx.pack_records(h_arena[:bases], offsets, 0, pk, fl)
        ctx.run_batch_packed(pk, fl, offsets, out=h_out2)
    torch.cuda.synchronize()
    e2e_packed_incl_s = time.perf_counter() - t0

    # results sanity: device-resident (Myers) and host (library default) legs agree, every planted region found (k = 0)
    dev_res = d_out.cpu().numpy().view(hx.RESULT_DTYPE).reshape(n, npairs)
    assert dev_res.tobytes() == h_out.tobytes(), "device-resident and host legs disagree"
    found = int((h_out["flags"] & 4).astype(bool).sum())
    import hashlib
    table_sha = hashlib.sha256(h_out.tobytes()).hexdigest()
    ctx.set_option("fast_small_k", 0)
    ctx.set_primers(pairs, args.k)  # group_info() below describes the graded Myers configuration

    # ---- one batch sharded over the N ranks (hyperex_b200/shard.py: contiguous record ranges, result tables gathered on
    # rank 0 in input order, no collective on the data path): equal to the same batch on one GPU ------------------------
    sharded = None
    if world > 1:
        from hyperex_b200 import shard
        gloo = dist.new_group(backend="gloo")  # the gather of the small result tables is host-side plumbing (CPU tensors)
        sa, so = hx_synth.gen_reads(24_000, seed=99)
        part = shard.run_sharded(lambda aa, oo: ctx.run_batch(aa, oo), sa, so, rank, world, npairs, hx.RESULT_DTYPE, group=gloo)
        if rank == 0:
            sharded = bool(part.tobytes() == ctx.run_batch(sa, so).tobytes())
        dist.barrier(group=gloo)

    # ---- one process, all N GPUs in one context (what `bin/hyperex --gpus N` ships): parity outside the timed region --
    multidev = None
    if world > 1:
        barrier()
        if rank == 0:
            ns = min(n, 20_000)
            so = np.ascontiguousarray(offsets[:ns + 1])
            with hx.Context(tuple(range(world))) as mctx:
                mctx.set_option("slab_bytes", 2 << 20)
                mctx.set_primers(pairs, args.k)
                mres = mctx.run_batch(arena[:int(so[-1])], so)
            multidev = bool(mres.tobytes() == h_out[:ns].tobytes())
        barrier()

    # ---- file-level leg: the drop-in seam (FASTA file in, two files out), rank 0, a context over all N GPUs ----------
    file_e2e = None
    if not args.no_file_e2e and args.workload == "c2":
        barrier()
        if rank == 0:
            try:
                file_e2e = file_e2e_leg(hx, range(world), arena, offsets, pairs, args.k, max(1, args.file_steps), args.file_dir)
            except Exception as e:  # pragma: no cover - reported, never hidden
                file_e2e = {"error": str(e)}
        barrier()

    # ---- max over ranks ---------------------------------------------------------------------------
    torch.cuda.set_device(local)
    dev_local = torch.device("cuda", local)
    t = torch.tensor([dev_ms, e2e_s * 1e3, myers_ms, fast_ms or 0.0, h2d_conc_ms, e2e_packed_s * 1e3, e2e_packed_incl_s * 1e3,
                      pack_s * 1e3, e2e_ascii_s * 1e3, e2e_otw_s * 1e3, e2e_hyb_s * 1e3], dtype=torch.float64, device=dev_local)
    tot = torch.tensor([float(bases), float(launches)], dtype=torch.float64, device=dev_local)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    (dev_ms_max, e2e_ms_max, myers_ms_max, fast_ms_max, h2d_conc_ms_max, e2e_pk_ms_max, e2e_pk_incl_ms_max,
     pack_ms_max, e2e_ascii_ms_max, e2e_otw_ms_max, e2e_hyb_ms_max) = (float(x) for x in t.cpu())
    total_bases, total_launches = (float(x) for x in tot.cpu())

    if rank == 0:
        value = total_bases * args.steps / (dev_ms_max * 1e-3) / 1e9
        e2e = total_bases * args.steps / (e2e_ms_max * 1e-3) / 1e9
        hbm_peak, hbm_src = load_peaks()
        # --- roofline of the dominant kernel (hx_myers_pair_kernel) -------------------------------
        # algorithmic work per launch: the launch scans every base of this rank
        # against the 15 unique built-in patterns (20 searches/base in the reference's count);
        # one word-step = INT_OPS_PER_STEP integer instructions (SASS of the hot loop).
        n_groups, steps32, steps64 = ctx.group_info()
        uniq = steps32 + steps64  # automata stepped per text byte (15 for the 10 built-in regions)
        # Instruction mix of the hot loop per word-step at NP = 15 (cuobjdump -sass + the ncu source page,
        # profiles/r01_ncu_myers_pair_kernel_u32_15.csv): ALU pipe 10.0 (7 LOP3 + 2 LEA.HI + 0.5 LOP3 hit test +
        # 0.5 per-byte overhead), FMA pipe 3.1 (IMAD.IADD: the carry add and the two shifts).
        alu_ops, fma_ops = 10.0, 3.1
        per_launch_ms = myers_ms_max / max(1, myers_launches)
        steps_exec = bases * uniq                       # word-steps of one pass over the batch (all groups)
        scan_ms = myers_ms_max / args.steps             # all scan launches of one step
        steps_per_s = steps_exec / (scan_ms * 1e-3)
        peak = {}
        for name, var in (("lop3", 0), ("imad", 3), ("lop3_imad", 4), ("myers_step", 10)):
            try:
                peak[name] = hx.int_peak(var, 1500, local) / 1e3  # Tops/s
            except Exception as e:  # pragma: no cover
                peak[name] = None
                peak["error"] = str(e)
        alu_peak = peak.get("lop3") or float("nan")
        achieved_alu = steps_per_s * alu_ops / 1e12
        roofline = {"bound": "int_alu", "kernel": "hx_myers_pair_kernel<u32,15>" if args.workload == "c2" else
                    f"hx_myers_pair_kernel x {n_groups} groups ({steps32} u32 + {steps64} u64 automata/base; the ALU "
                    "op counts below are those of the u32 NP=15 kernel)",
                    "achieved": achieved_alu, "peak": alu_peak, "unit": "Tint-op/s (ALU pipe)",
                    "frac": achieved_alu / alu_peak,
                    "peak_source": "of measured: hx_int_peak LOP3 stream in this run (the ALU pipe executes LOP3/LEA/"
                                   "SHF/IADD3; MEASURED_PEAKS.json has no integer figure)",
                    "word_steps_per_s": steps_per_s, "alu_ops_per_word_step": alu_ops,
                    "fma_ops_per_word_step": fma_ops, "kernel_ms_per_launch": per_launch_ms,
                    "kernel_share_of_step": myers_ms_max / dev_ms_max,
                    "all_int_ops": {"achieved": steps_per_s * (alu_ops + fma_ops) / 1e12,
                                    "dual_pipe_peak": peak.get("lop3_imad"),
                                    "frac": (steps_per_s * (alu_ops + fma_ops) / 1e12 / peak["lop3_imad"])
                                    if peak.get("lop3_imad") else None},
                    "bare_step_peak_gsteps": (peak.get("myers_step") or 0) * 1e3,
                    "traffic": None, "traffic_detail": load_traffic(bases) if args.workload == "c2" else None,
                    "hbm": {"achieved": bases * n_groups / (scan_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": bases * n_groups / (scan_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                            "algorithmic_bytes_per_launch": bases}}
        if roofline["traffic_detail"]:  # contract: `traffic` = dram bytes read + written per launch (ncu), a number
            roofline["traffic"] = roofline["traffic_detail"]["bytes"]
        if ab_ms is not None:
            roofline["text_path_ab"] = {"default_ms_per_step": dev_ms_max / args.steps, "tma_ms_per_step": ab_ms,
                                        "what": "default: per-thread 16-byte cp.async copies into shared memory, one chunk ahead; "
                                                "tma: per-thread cp.async.bulk + mbarrier (option text_tma)"}
        line = {"metric": "Gbases/s", "value": value, "unit": "Gbases/s", "gcups": value * cells,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": workload_desc, "records_per_gpu": n,
                           "automata_per_base": {"u32": steps32, "u64": steps64, "groups": n_groups,
                                                 "long_filter": args.long_filter},
                           "bases_per_gpu": bases, "pairs": npairs, "k": args.k, "regions_found": found,
                           "l2": f"inputs ({bases / 1e9:.2f} GB/GPU) larger than the 126 MB L2; no flush needed",
                           "parallelism": f"records sharded over {world} GPU(s), no collective",
                           "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
                "e2e": {"value": e2e, "unit": "Gbases/s", "gcups": e2e * cells, "ms_per_step": e2e_ms_max / args.steps,
                        "h2d_bytes_per_step": e2e_stats["h2d_bytes"], "d2h_bytes_per_step": e2e_stats["d2h_bytes"],
                        "timer": "host wall clock around hx_run_batch (synchronous), max over ranks",
                        "packed_on_the_way": e2e_stats["packed_text"], "mode": e2e_stats["mode"],
                        "slabs_packed": e2e_stats["slabs_packed"], "pack_threads": e2e_stats["pack_threads"],
                        "slabs": e2e_stats["slabs"],
                        "ascii_slabs": {"value": total_bases * args.steps / (e2e_ascii_ms_max * 1e-3) / 1e9, "unit": "Gbases/s",
                                        "ms_per_step": e2e_ascii_ms_max / args.steps, "h2d_bytes_per_step": ascii_stats["h2d_bytes"],
                                        "what": "pack_h2d = 0: the arena crosses PCIe as it is, one byte per base"},
                        "packed_slabs": {"value": total_bases * args.steps / (e2e_otw_ms_max * 1e-3) / 1e9, "unit": "Gbases/s",
                                         "ms_per_step": e2e_otw_ms_max / args.steps, "h2d_bytes_per_step": otw_stats["h2d_bytes"],
                                         "pack_threads": otw_stats["pack_threads"],
                                         "what": "pack_h2d = 1: every slab is packed to 4 bits per base by the host threads (inside the "
                                                 "timed call) while the slabs before it are copied and scanned"},
                        "mixed_slabs": {"value": total_bases * args.steps / (e2e_hyb_ms_max * 1e-3) / 1e9, "unit": "Gbases/s",
                                        "ms_per_step": e2e_hyb_ms_max / args.steps, "h2d_bytes_per_step": hyb_stats["h2d_bytes"],
                                        "pack_threads": hyb_stats["pack_threads"], "slabs": hyb_stats["slabs"],
                                        "slabs_packed": hyb_stats["slabs_packed"],
                                        "what": "pack_h2d = 2: ASCII slabs from the front of the batch keep the link busy while the "
                                                "host threads pack slabs from its back; the two meet where this host makes them meet "
                                                "(rank 0's last call shown)"},
                        "h2d_copy_alone_ms": h2d_only_ms,
                        "h2d_copy_alone_gbs": bases / (h2d_only_ms * 1e-3) / 1e9,
                        "frac_of_copy_alone": h2d_only_ms / (e2e_ms_max / args.steps),
                        "copy_alone_note": "one plain pinned -> device copy of the ASCII arena; a packed path can beat it (half the bytes)",
                        "h2d_copy_all_ranks_at_once_ms": h2d_conc_ms_max,
                        "h2d_copy_all_ranks_at_once_gbs_total": total_bases / (h2d_conc_ms_max * 1e-3) / 1e9,
                        "frac_of_copy_all_ranks_at_once": h2d_conc_ms_max / (e2e_ms_max / args.steps)},
                "gpu_launches": int(total_launches), "roofline": roofline, "clocks": clocks,
                "result_table_sha256_rank0": table_sha}
        line["e2e"]["path"] = ("library defaults (fast_small_k = 1, pack_h2d automatic; --pack-h2d %d given): result table asserted "
                               "equal to the Myers kernel's and equal between the ASCII, packed and mixed paths" % args.pack_h2d)
        line["e2e_packed"] = {
            "what": "hx_run_batch_packed: pinned PACKED arena (4-bit codes, hx_pack_records) -> H2D -> kernels on nibble text -> D2H "
                    "result table, every step; result table asserted equal to the ASCII leg's",
            "value": total_bases * args.steps / (e2e_pk_ms_max * 1e-3) / 1e9, "unit": "Gbases/s",
            "ms_per_step": e2e_pk_ms_max / args.steps, "h2d_bytes_per_step": int(bases // 2 + offsets.nbytes + n),
            "d2h_bytes_per_step": int(n * npairs * 12), "pack_inside_timed_region": False,
            "pack_ms": pack_ms_max, "pack_gbases_per_s": bases / (pack_ms_max * 1e-3) / 1e9, "pack_threads": os.cpu_count(),
            "with_pack_inside_timed_region": {"value": total_bases * args.steps / (e2e_pk_incl_ms_max * 1e-3) / 1e9,
                                              "ms_per_step": e2e_pk_incl_ms_max / args.steps,
                                              "note": "pack then run, back to back on the host (not pipelined)"}}
        if sharded is not None:
            line["sharded_parity"] = sharded
        if multidev is not None:
            line["multidev_parity"] = multidev
        if file_e2e is not None:
            line["file_e2e"] = file_e2e
        if fast_ms_max > 0:
            fv = total_bases / (fast_ms_max * 1e-3) / 1e9
            line["fast_path"] = {"what": "library default for -m 0..2: Shift-Or / Wu-Manber automata of primer suffixes, four <= 8-symbol "
                                         "automata per word at -m 0 and two <= 16-symbol ones at -m 1..2, + verification (DESIGN.md "
                                         "3.5, 3.8), identical results (asserted), not counted as Myers GCUPS; ms_per_step is the "
                                         "whole device step (tile build, classify, scan, combine)",
                                 "value": fv, "unit": "Gbases/s",
                                 "ms_per_step": fast_ms_max, "speedup_vs_myers": dev_ms_max / args.steps / fast_ms_max,
                                 "groups": fast_groups[0], "automata_per_base": fast_groups[1] + fast_groups[2]}
            # per text byte and thread, packed kernel (4 words at -m 0, 8 at -m 1..2; SASS of the hot loop / ncu source
            # page, profiles/r02_sass_*, r02_hot_*): ALU-pipe instructions, FMA-pipe instructions (IMAD.IADD shifts),
            # LDS.128 of Peq rows
            mix = {0: (11.4, 5.8, 1.0), 1: (40.0, 21.0, 2.0), 2: (64.0, 35.0, 2.0)}.get(args.k)
            if mix and args.workload == "c2":
                sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
                lds_peak = 148 * sm_hz * 128 / (mix[2] * 16)  # bases/s the 128 B/clk/SM shared-memory pipe allows
                issue_peak = 148 * 4 * sm_hz * 32 / (mix[0] + mix[1] + mix[2] + 1.0)
                line["fast_path"]["roofline"] = {
                    "alu_ops_per_base": mix[0], "fma_ops_per_base": mix[1], "lds128_per_base": mix[2],
                    "alu": {"achieved": fv * mix[0] / 1e3, "peak": alu_peak, "unit": "Tint-op/s (ALU pipe)",
                            "frac": fv * mix[0] / 1e3 / alu_peak},
                    "lds": {"achieved": fv, "peak": lds_peak / 1e9, "unit": "Gbases/s at 128 B/clk/SM", "frac": fv * 1e9 / lds_peak},
                    "issue": {"achieved": fv, "peak": issue_peak / 1e9, "unit": "Gbases/s at 1 warp-instruction/clk/SMSP",
                              "frac": fv * 1e9 / issue_peak},
                    "note": "no pipe is saturated: with 15 primer sites in every read a third of all 4-byte trips is replayed "
                            "symbol by symbol by the lane that saw a hit and the drain verifies the rest of the primer "
                            "(hot loop = 67 % of the executed instructions, 46 % of the stall samples; "
                            "profiles/r02_ncu_shiftor_k0_quarters.csv); the fractions include the tile-build / classify / "
                            "combine kernels of the step in the time"}
        if not args.no_cpu_baseline:
            ns = min(args.cpu_sample, n)
            cb, cdt = cpu_port_run(arena, offsets, ns, pairs, args.k, 1)
            line["cpu_baseline"] = {"value": cb / cdt / 1e9, "unit": "Gbases/s", "gcups": cb / cdt / 1e9 * cells,
                                    "cores": 1, "kind": "port",
                                    "sample": f"first {ns} records ({cb} bases) of the same workload, 1 thread (the "
                                              "reference is single-threaded); C port of utils.rs:270-351 + rust-bio "
                                              f"Myers, {cdt:.1f} s"}
        print(json.dumps(line), flush=True)
    hx.pinned_free(h_arena)
    hx.pinned_free(h_out_raw)
    for buf in (pk, fl, h_out2_raw):
        hx.pinned_free(buf)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
