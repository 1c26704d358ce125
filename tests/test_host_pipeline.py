"""CPU suite for the host pipeline of hx_get_hypervar_regions (hyperex_b200/csrc/hx_host.cpp): chunker, parallel
parsers, lanes that finish out of order, ordered reservation of the output byte ranges, ordered log lines, the pwrite
pool and the error paths.  The device runtime behind the internal lane interface is replaced by tests/stub/
hx_lane_stub.cpp, whose "lanes" compute with the CPU oracle and sleep for random times; everything in front of and
behind the lanes is the product's own hx_host.cpp, compiled unchanged.  The same pipeline is byte-compared on the GPU
by tests/test_gpu_cli.py and tests/test_gpu_parity.py."""
import bz2
import ctypes as C
import gzip
import lzma
import os
import subprocess

import numpy as np
import pytest

import hx_synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG_FN = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)


@pytest.fixture(scope="module")
def stub():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhxhost_stub.so")
    srcs = [os.path.join(ROOT, "tests", "stub", "hx_lane_stub.cpp"), os.path.join(ROOT, "hyperex_b200", "csrc", "hx_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "hyperex_b200", "csrc", "hx_host.h"), os.path.join(ROOT, "include", "hyperex_b200.h"),
                   os.path.join(ROOT, "oracle", "hx_oracle.c")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        obj = os.path.join(out, "hx_oracle.o")
        subprocess.run(["gcc", "-O2", "-fPIC", "-std=c11", "-pthread", "-c", os.path.join(ROOT, "oracle", "hx_oracle.c"), "-o", obj],
                       check=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I/usr/local/cuda/include", "-o", so] + srcs +
                       [obj, "-lz", "-ldl", "-lpthread"], check=True)
    L = C.CDLL(so)
    L.stub_ctx_new.restype = C.c_void_p
    L.stub_ctx_new.argtypes = [C.c_int, C.c_int]
    L.stub_ctx_free.argtypes = [C.c_void_p]
    L.hx_get_hypervar_regions.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                          C.c_char_p, C.c_uint8, LOG_FN, C.c_void_p]
    L.hx_file_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    L.hx_last_error.restype = C.c_char_p
    L.hx_last_error.argtypes = [C.c_void_p]
    return L


def run(stub, path, pairs, prefix, k, lanes=4, fail_at=-1, log=True, env=None):
    """-> (rc, fa bytes, gff bytes, [(level, msg)], stats)"""
    ctx = stub.stub_ctx_new(lanes, fail_at)
    lines = []
    cb = LOG_FN((lambda lvl, msg, _u: lines.append((lvl, msg.decode()))) if log else 0)
    fw = (C.c_char_p * len(pairs))(*[p[0].encode() for p in pairs])
    rv = (C.c_char_p * len(pairs))(*[p[1].encode() for p in pairs])
    old = {}
    for key, val in (env or {}).items():
        old[key] = os.environ.get(key)
        os.environ[key] = str(val)
    try:
        rc = stub.hx_get_hypervar_regions(ctx, str(path).encode(), len(pairs), fw, rv, str(prefix).encode(), k, cb, None)
    finally:
        for key, val in old.items():
            if val is None:
                del os.environ[key]
            else:
                os.environ[key] = val
    stats = (C.c_double * 14)()
    stub.hx_file_stats(ctx, stats, 14)
    err = stub.hx_last_error(ctx).decode()
    stub.stub_ctx_free(ctx)
    fa = open(str(prefix) + ".fa", "rb").read() if os.path.exists(str(prefix) + ".fa") else None
    gff = open(str(prefix) + ".gff", "rb").read() if os.path.exists(str(prefix) + ".gff") else None
    return rc, fa, gff, lines, list(stats), err


def expected_logs(arena, offsets, ids, pairs, res):
    """the reference's warn!/error! lines (utils.rs:276-288, 387-408) from the oracle's result table"""
    out = []
    labels = [O.primers_to_region(f, r) for f, r in pairs]
    for r, rid in enumerate(ids):
        n = int(offsets[r + 1] - offsets[r])
        if res[r, 0]["flags"] & 8:
            out.append((2, f"Unrecognized sequence type in record: {rid}"))
            continue
        if n <= 1500:
            out.append((1, f"Sequence {rid} is short ({n} bp). Some regions may not be found."))
        for p, (f, rv) in enumerate(pairs):
            fl = int(res[r, p]["flags"])
            fa, ra, ok = fl & 1, fl & 2, fl & 4
            if fa and ra and ok:
                continue
            if fa and ra:
                out.append((1, f"Forward and reverse primers for region {labels[p]} were both found in sequence {rid}, "
                               "but not in a valid order/orientation to define a region"))
            elif not fa and ra:
                out.append((1, f"Forward primer {f} not found for region {labels[p]} in sequence {rid}"))
            elif fa and not ra:
                out.append((1, f"Reverse primer {rv} not found for region {labels[p]} in sequence {rid}"))
            else:
                out.append((1, f"Both primers not found for region {labels[p]} in sequence {rid}"))
    return out


def make_input(n=400, seed=11, **kw):
    a, o = hx_synth.gen_reads_mutated(n, seed=seed, max_edits=3, revcomp_frac=0.3, lower_frac=0.3, rna_frac=0.1, **kw)
    ids = [f"rec{i}" if i % 4 else f"rec{i} with a description" for i in range(n)]
    return hx_synth.to_fasta(a, o, ids, width=70, crlf=False)


@pytest.mark.parametrize("codec", ["plain", "gz", "bz2", "xz"])
@pytest.mark.parametrize("host_format", [0, 1])
def test_pipeline_bytes_and_logs_equal_the_oracle(stub, tmp_path, codec, host_format):
    buf = make_input()
    pairs = O.default_pairs()
    path = tmp_path / f"in.fa.{codec}"
    path.write_bytes({"plain": lambda b: b, "gz": gzip.compress, "bz2": bz2.compress, "xz": lzma.compress}[codec](buf))
    efa, egff, eres = O.get_hypervar_regions(buf, pairs, 2)
    arena, offsets, ids = O.parse_fasta(buf)
    elog = expected_logs(arena, offsets, ids, pairs, eres)
    assert len(efa) > 10000
    for chunk in (3000, 50000, 0):  # ~200 batches / ~12 batches / one batch
        env = {"HYPEREX_HOST_FORMAT": host_format}
        if chunk:
            env["HYPEREX_CHUNK_BYTES"] = chunk
        rc, fa, gff, lines, stats, err = run(stub, path, pairs, tmp_path / f"out{chunk}", 2, env=env)
        assert rc == 0, err
        assert fa == efa and gff == egff
        assert lines == elog
        assert stats[1] == int(offsets[-1]) and stats[2] == len(ids)          # bases, records
        assert stats[9] == len(efa) and stats[10] == len(egff)
        if chunk == 3000:
            assert stats[3] > (50 if codec == "plain" else 5)                 # really many batches


def test_pipeline_without_log_fn_and_single_lane(stub, tmp_path):
    buf = make_input(150, seed=3)
    pairs = [O.region_to_primer("v3v4"), O.region_to_primer("v4v5")]
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    efa, egff, _ = O.get_hypervar_regions(buf, pairs, 1)
    for lanes in (1, 2, 7):
        for wmode in ("mmap", "pwrite"):  # both ways the write pool moves bytes into the files
            rc, fa, gff, lines, _s, err = run(stub, path, pairs, tmp_path / f"o{lanes}{wmode}", 1, lanes=lanes, log=False,
                                              env={"HYPEREX_CHUNK_BYTES": 5000, "HYPEREX_WRITE": wmode})
            assert rc == 0 and fa == efa and gff == egff and lines == [], err


def test_large_text_blocks_are_written_in_parallel_slices(stub, tmp_path):
    """one batch whose text is several MB: the write pool cuts it into page-aligned slices written by different threads"""
    a, o = hx_synth.gen_reads(1500, seed=21)
    buf = hx_synth.to_fasta(a, o)
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    pairs = O.default_pairs()
    efa, egff, _ = O.get_hypervar_regions(buf, pairs, 0)
    assert len(efa) > (8 << 20)
    (tmp_path / "om.gff").write_bytes(b"x" * 1234)  # append at an odd offset: slices are aligned to FILE pages
    for wmode, name in (("mmap", "om"), ("pwrite", "op")):
        rc, fa, gff, _l, _s, err = run(stub, path, pairs, tmp_path / name, 0, lanes=2, log=False,
                                          env={"HYPEREX_WRITE": wmode, "HYPEREX_HOST_FORMAT": 1})  # one write call per batch and file
        assert rc == 0 and fa == efa, err
        assert gff == (b"x" * 1234 if name == "om" else b"") + egff


def test_pipeline_stops_where_the_reference_reader_stops(stub, tmp_path):
    """a malformed / empty record ends rust-bio's iteration: nothing after it is processed, whatever chunk it is in"""
    good = make_input(120, seed=5)
    pairs = O.default_pairs()
    for breaker in (b">\n", b"garbage line\n"):
        # "garbage line" directly after a record would be a sequence line; put it where a header is expected
        buf = good + (b">\n" if breaker == b">\n" else b"") + make_input(60, seed=6)
        if breaker != b">\n":
            buf = b"not a header\n" + good
        path = tmp_path / "in.fa"
        path.write_bytes(buf)
        efa, egff, eres = O.get_hypervar_regions(buf, pairs, 0)
        for chunk in (2000, 0):
            rc, fa, gff, lines, stats, err = run(stub, path, pairs, tmp_path / f"o{len(breaker)}_{chunk}", 0,
                                                     env={"HYPEREX_CHUNK_BYTES": chunk} if chunk else None)
            assert rc == 0, err
            assert fa == efa and gff == egff
            assert stats[2] == eres.shape[0]


def test_pipeline_error_paths_do_not_hang(stub, tmp_path):
    buf = make_input(200, seed=8)
    pairs = O.default_pairs()
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    for host_format in (0, 1):
        for fail_at in (0, 3, 17):
            rc, _fa, _gff, _lines, _s, err = run(stub, path, pairs, tmp_path / "o", 0, fail_at=fail_at,
                                                 env={"HYPEREX_CHUNK_BYTES": 4000, "HYPEREX_HOST_FORMAT": host_format})
            assert rc == -2 and "injected" in err
    # unreadable input / output
    rc, *_ = run(stub, tmp_path / "missing.fa", pairs, tmp_path / "o", 0)
    assert rc == -5
    rc, *_ = run(stub, path, pairs, tmp_path / "no_such_dir" / "o", 0)
    assert rc == -5
    # a truncated gzip stream: the records before the damage are written, the call still succeeds (rust-bio's
    # iterator ends at the failing read; `.flatten()` hides the error, utils.rs:270)
    gz = gzip.compress(buf)
    bad = tmp_path / "bad.fa.gz"
    bad.write_bytes(gz[: len(gz) // 2])
    rc, fa, gff, _l, stats, err = run(stub, bad, pairs, tmp_path / "ob", 0, env={"HYPEREX_CHUNK_BYTES": 4000})
    assert rc == 0, err
    efa, _egff, _ = O.get_hypervar_regions(buf, pairs, 0)
    assert 0 < stats[2] < 200 and efa.startswith(fa) and len(fa) > 0


def test_gff_is_appended_like_the_reference(stub, tmp_path):
    """utils.rs:242-245 opens <prefix>.gff with create+append; <prefix>.fa is truncated (utils.rs:241)"""
    buf = make_input(30, seed=9)
    pairs = O.default_pairs()
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    (tmp_path / "o.gff").write_bytes(b"previous content\n")
    (tmp_path / "o.fa").write_bytes(b"x" * 100000)
    efa, egff, _ = O.get_hypervar_regions(buf, pairs, 0)
    rc, fa, gff, *_ = run(stub, path, pairs, tmp_path / "o", 0)
    assert rc == 0 and fa == efa and gff == b"previous content\n" + egff


def test_chunk_boundaries_are_record_starts(stub, tmp_path):
    """'>' inside a header or a sequence line is not a record start; only a line that begins with '>' is"""
    recs = []
    rng = np.random.default_rng(4)
    for i in range(300):
        seq = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 200))))
        if i % 3 == 0:
            seq = seq[: len(seq) // 2] + ">" + seq[len(seq) // 2:]
        recs.append(f">id{i}>x desc>y\n{seq}\n" if i % 5 else f">id{i}\r\n{seq}\r\n\r\n")
    buf = "".join(recs).encode()
    path = tmp_path / "in.fa"
    path.write_bytes(buf)
    pairs = [["ACGT", "GGCC"]]  # README.md:98: k >= a primer's length -> every position is a hit
    efa, egff, eres = O.get_hypervar_regions(buf, pairs, 1)
    for chunk in (64, 777, 0):
        rc, fa, gff, _l, stats, err = run(stub, path, pairs, tmp_path / f"o{chunk}", 1, env={"HYPEREX_CHUNK_BYTES": chunk} if chunk else None)
        assert rc == 0 and fa == efa and gff == egff, err
        assert stats[2] == 300
