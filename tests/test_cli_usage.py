"""CPU suite for the command line (`bin/hyperex`, src/cli.rs + src/main.rs): everything that is decided before a
device is needed — flag conflicts, value parsing, region / primer validation, refusal to overwrite — and that the
binary fails loudly, without writing outputs, when no GPU is usable (there is no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "hyperex")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="bin/hyperex not built (run __graft_entry__.build())")


def run(args, cwd, stdin=None):
    return subprocess.run([BIN] + args, cwd=cwd, input=stdin, capture_output=True, timeout=120)


def test_flag_conflicts_and_value_parsing(tmp_path, golden_dir):
    fa = os.path.join(golden_dir, "test.fa")
    assert run([fa, "-f", "ACGT"], tmp_path).returncode == 2                                   # cli.rs: requires = "reverse"
    assert run([fa, "-r", "ACGT"], tmp_path).returncode == 2                                   # requires = "forward"
    assert run([fa, "-f", "ACGT", "-r", "ACGT", "--region", "v4"], tmp_path).returncode == 2   # conflicts_with
    assert run([fa, "-m", "300"], tmp_path).returncode == 2                                    # not a u8
    assert run([fa, "-m", "-1"], tmp_path).returncode == 2
    assert run([fa, "--no-such-flag"], tmp_path).returncode == 2
    assert run(["--version"], tmp_path).stdout.strip() == b"hyperex 0.2.0"
    assert run(["--help"], tmp_path).returncode == 0
    assert not any(p.suffix in (".fa", ".gff") for p in tmp_path.iterdir())


def test_input_and_region_validation(tmp_path, golden_dir):
    fa = os.path.join(golden_dir, "test.fa")
    assert run(["/nonexistent.fa"], tmp_path).returncode == 1
    r = run([fa, "--region", "v9v9"], tmp_path)
    assert r.returncode == 1 and b"not a correct file name nor a supported region name" in r.stderr
    r = run([fa, "-m", "30", "-p", "big"], tmp_path)                                           # utils.rs:522-540
    assert r.returncode == 1 and b"greater than length of longest primer" in r.stdout
    assert not (tmp_path / "big.fa").exists()


def test_refuses_to_overwrite_without_force(tmp_path, golden_dir):
    fa = os.path.join(golden_dir, "test.fa")
    (tmp_path / "keep.fa").write_bytes(b"precious")
    r = run([fa, "-p", "keep"], tmp_path)                                                      # utils.rs:455-481
    assert r.returncode == 1 and b"already exists" in r.stderr
    assert (tmp_path / "keep.fa").read_bytes() == b"precious"


def test_no_gpu_is_an_error_not_a_fallback(tmp_path, golden_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run([os.path.join(golden_dir, "test.fa"), "-p", "out"], tmp_path)
    assert r.returncode == 1 and b"no CUDA device" in r.stderr and b"no CPU fallback" in r.stderr
    assert not (tmp_path / "out.fa").exists() and not (tmp_path / "out.gff").exists()
