"""Times the drop-in CLI end to end (file in, files out) on a synthetic FASTA; prints one JSON line."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import hx_synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
gpus = sys.argv[2] if len(sys.argv) > 2 else "1"
import shutil  # noqa: E402

# tmpfs when it has room (the run measures the program, not a disk); HX_CLI_DIR overrides
base = os.environ.get("HX_CLI_DIR") or ("/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 12 * n * 1600
                                        else os.path.join(ROOT, "gpurun_out"))
work = os.path.join(base, "cli_e2e")
os.makedirs(work, exist_ok=True)
a, o = hx_synth.gen_reads(n, seed=2)
path = os.path.join(work, "in.fa")
with open(path, "wb") as f:
    chunk = []
    for i in range(n):
        chunk.append(b">r%d\n" % i + a[int(o[i]):int(o[i + 1])].tobytes() + b"\n")
        if len(chunk) == 20000:
            f.write(b"".join(chunk))
            chunk = []
    f.write(b"".join(chunk))
t0 = time.perf_counter()
r = subprocess.run([os.path.join(ROOT, "bin", "hyperex"), path, "-p", os.path.join(work, "out"), "--force", "-q",
                    "--gpus", gpus], cwd=work, capture_output=True, env=dict(os.environ, HYPEREX_TIMING=os.environ.get("HYPEREX_TIMING", "1")))
dt = time.perf_counter() - t0
out = {"records": n, "bases": int(o[-1]), "gpus": gpus, "wall_s": dt, "gbases_per_s": int(o[-1]) / dt / 1e9,
       "dir": base, "rc": r.returncode, "timing": r.stderr.decode()[-3000:].strip(),
       "fa_bytes": os.path.getsize(os.path.join(work, "out.fa")), "gff_bytes": os.path.getsize(os.path.join(work, "out.gff"))}
shutil.rmtree(work, ignore_errors=True)  # gpurun_out/ must stay small
print(json.dumps(out))
