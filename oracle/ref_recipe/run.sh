#!/bin/bash
# Generates reference-made golden vectors into tests/golden/ref/ (needs cargo + network; NOT runnable in the build image).
#   1. find_all_lazy tables from rust-bio 1.6.0 for tests/golden/ref_cases.tsv
#   2. FASTA / GFF3 bytes from the unmodified hyperex v0.2.0 binary for the committed fixtures
# usage: oracle/ref_recipe/run.sh [path to a hyperex v0.2.0 checkout]   (default: cargo install hyperex --version 0.2.0)
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$ROOT/tests/golden/ref"
mkdir -p "$OUT"
( cd "$ROOT/oracle/ref_recipe" && cargo build --release )
"$ROOT/oracle/ref_recipe/target/release/hx_ref_vectors" < "$ROOT/tests/golden/ref_cases.tsv" > "$OUT/find_all_lazy.jsonl"
if [ $# -ge 1 ]; then ( cd "$1" && cargo build --release ); HX="$1/target/release/hyperex"; else cargo install hyperex --version 0.2.0 --root "$OUT/.cargo"; HX="$OUT/.cargo/bin/hyperex"; fi
W="$(mktemp -d)"
run() { # name, input, args...
  local name="$1" input="$2"; shift 2
  ( cd "$W" && "$HX" "$input" -p "$W/$name" --force -q "$@" > /dev/null ) || true
  cp "$W/$name.fa" "$OUT/$name.fa"; cp "$W/$name.gff" "$OUT/$name.gff"
}
G="$ROOT/tests/golden"
run config1_k0 "$G/test.fa"
run config1_k2 "$G/test.fa.gz" -m 2
run primers_csv_k3 "$G/test.fa" --region "$G/primers.txt" -m 3
for f in "$G"/ref_inputs/*.fa; do b="$(basename "$f" .fa)"; run "${b}_k0" "$f"; run "${b}_k2" "$f" -m 2; run "${b}_v3v4_v4v5_k3" "$f" --region v3v4 v4v5 -m 3; done
( cd "$OUT" && sha256sum *.fa *.gff *.jsonl > SHA256SUMS )
echo "wrote $(ls "$OUT" | wc -l) files to $OUT (provenance: $("$HX" --version), rust-bio 1.6.0, $(date -u +%F))" | tee "$OUT/PROVENANCE.txt"
