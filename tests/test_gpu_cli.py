"""GPU suite for the drop-in command line (`bin/hyperex`, src/cli.rs + src/main.rs): same flags, same FASTA and
GFF3 bytes as the reference path (checked against the oracle / the committed goldens)."""
import gzip
import os
import subprocess

import pytest

import hx_synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "hyperex")


def run(args, cwd, stdin=None, env=None):
    return subprocess.run([BIN] + args, cwd=cwd, input=stdin, capture_output=True, timeout=300,
                          env=dict(os.environ, **(env or {})))


def test_default_regions_on_reference_fixture(tmp_path, golden_dir):
    r = run([os.path.join(golden_dir, "test.fa")], tmp_path)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "hyperex_out.fa").read_bytes() == open(os.path.join(golden_dir, "config1_k0.fa"), "rb").read()
    assert (tmp_path / "hyperex_out.gff").read_bytes() == open(os.path.join(golden_dir, "config1_k0.gff"), "rb").read()
    out = r.stdout.decode()
    assert "This is hyperex v0.2.0" in out and "Both primers not found for region v1v2" in out
    assert "Walltime:" in out and "Enjoy. Share. Come back again!" in out
    assert (tmp_path / "hyperex.log").exists()
    # outputs exist now: refused without --force (utils.rs:455-481), exit code 1
    r2 = run([os.path.join(golden_dir, "test.fa")], tmp_path)
    assert r2.returncode == 1 and b"already exists" in r2.stderr
    r3 = run([os.path.join(golden_dir, "test.fa.gz"), "--force", "-q", "-m", "2"], tmp_path)
    assert r3.returncode == 0 and b"[INFO]" not in r3.stdout.replace(b"\x1b[32m", b"[").replace(b"\x1b[0m", b"]")
    assert (tmp_path / "hyperex_out.gff").read_bytes() == open(os.path.join(golden_dir, "config1_k2.gff"), "rb").read()


def test_flags_region_primers_csv_prefix_stdin(tmp_path, golden_dir):
    a, o = hx_synth.gen_reads_mutated(120, seed=77, max_edits=2, revcomp_frac=0.3, lower_frac=0.3)
    fasta = hx_synth.to_fasta(a, o, width=80)
    (tmp_path / "in.fa").write_bytes(fasta)
    (tmp_path / "in.fa.gz").write_bytes(gzip.compress(fasta))
    # --region with names, -m, -p
    # num_args = 1.. (cli.rs:51): a file name directly after the region names is taken as a region
    r = run(["--region", "v3v4", "in.fa"], tmp_path)
    assert r.returncode == 1 and b"not a correct file name nor a supported region name" in r.stderr
    r = run(["--region", "v3v4", "v4v5", "-m", "2", "-p", "sub", "in.fa.gz"], tmp_path)
    assert r.returncode == 0, r.stderr
    pairs = [O.region_to_primer("v3v4"), O.region_to_primer("v4v5")]
    fa, gff, _ = O.get_hypervar_regions(fasta, pairs, 2)
    assert (tmp_path / "sub.fa").read_bytes() == fa and (tmp_path / "sub.gff").read_bytes() == gff
    # -f / -r
    r = run(["in.fa", "-f", "CCTACGGGNGGCWGCAG", "-r", "GACTACHVGGGTATCTAATCC", "--mismatch", "1", "--prefix", "fr"], tmp_path)
    assert r.returncode == 0, r.stderr
    fa, gff, _ = O.get_hypervar_regions(fasta, [pairs[0]], 1)
    assert (tmp_path / "fr.fa").read_bytes() == fa and (tmp_path / "fr.gff").read_bytes() == gff
    # primer CSV through --region (README usage; utils.rs:494-495)
    r = run(["in.fa", "--region", os.path.join(golden_dir, "primers.txt"), "-p", "csv", "-m", "3"], tmp_path)
    assert r.returncode == 0, r.stderr
    csv_pairs = [["CCTACGGGNGGCWGCAG", "ATTACCGCGGCTGCTGG"], ["GTGCCAGCMGCCGCGGTAA", "GACTACHVGGGTATCTAATCC"]]
    fa, gff, _ = O.get_hypervar_regions(fasta, csv_pairs, 3)
    assert (tmp_path / "csv.fa").read_bytes() == fa and (tmp_path / "csv.gff").read_bytes() == gff
    # stdin ("-"), plain FASTA only (utils.rs:444-453)
    r = run(["-", "-p", "fromstdin", "-q"], tmp_path, stdin=fasta)
    assert r.returncode == 0, r.stderr
    fa, gff, _ = O.get_hypervar_regions(fasta, O.default_pairs(), 0)
    assert (tmp_path / "fromstdin.fa").read_bytes() == fa and (tmp_path / "fromstdin.gff").read_bytes() == gff
    assert not (tmp_path / "infile.fa").exists()


def test_usage_errors(tmp_path, golden_dir):
    fa = os.path.join(golden_dir, "test.fa")
    assert run([fa, "-f", "ACGT"], tmp_path).returncode == 2                          # requires = "reverse"
    assert run([fa, "-f", "ACGT", "-r", "ACGT", "--region", "v4"], tmp_path).returncode == 2  # conflicts_with
    assert run([fa, "-m", "300"], tmp_path).returncode == 2                            # not a u8
    assert run(["/nonexistent.fa"], tmp_path).returncode == 1
    r = run([fa, "--region", "v9v9"], tmp_path)
    assert r.returncode == 1 and b"not a correct file name nor a supported region name" in r.stderr
    r = run([fa, "-m", "30", "-p", "big"], tmp_path)                                  # utils.rs:522-540
    assert r.returncode == 1 and b"greater than length of longest primer" in r.stdout
    assert run(["--version"], tmp_path).stdout.strip() == b"hyperex 0.2.0"


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("gpus,host_format", [(1, 1), (2, 0), (2, 1)])
def test_cli_host_formatter_and_gpus_flag(tmp_path, gpus, host_format):
    """the two byte-writing branches (on-device text, HYPEREX_HOST_FORMAT=1) and `--gpus N` give the bytes of the
    1-GPU default run, which test_flags_region_primers_csv_prefix_stdin compares with the oracle"""
    if gpus > _n_gpus():
        pytest.skip("needs two GPUs")
    a, o = hx_synth.gen_reads_mutated(2000, seed=78, max_edits=2, revcomp_frac=0.3, lower_frac=0.3, rna_frac=0.1)
    fasta = hx_synth.to_fasta(a, o, width=80)
    (tmp_path / "in.fa").write_bytes(fasta)
    env = {"HYPEREX_HOST_FORMAT": str(host_format), "HYPEREX_CHUNK_BYTES": "100000"}
    r = run(["in.fa", "-m", "2", "-p", "x", "--gpus", str(gpus)], tmp_path, env=env)
    assert r.returncode == 0, r.stderr
    fa, gff, _ = O.get_hypervar_regions(fasta, O.default_pairs(), 2)
    assert (tmp_path / "x.fa").read_bytes() == fa and (tmp_path / "x.gff").read_bytes() == gff
    r1 = run(["in.fa", "-m", "2", "-p", "y"], tmp_path)
    strip = lambda b: [ln.split(b"] ", 1)[-1] for ln in b.splitlines() if b"WARN" in ln or b"ERROR" in ln]
    assert strip(r.stdout) == strip(r1.stdout) and len(strip(r.stdout)) > 100
    assert run(["in.fa", "--gpus", "0"], tmp_path).returncode == 2 and run(["in.fa", "--gpus", "x"], tmp_path).returncode == 2
