"""Hot-loop SASS extract of one instantiation of the scan kernel, from the built library (no GPU needed):
   python tools/sass_extract.py '<mangled kernel name>' [out.txt]
Finds the smallest backward-branch loop that holds the shared-memory Peq loads (LDS.128), prints its instructions with the
scheduling control bits (stall count, write / read barrier, wait mask) and the per-class instruction counts."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hyperex_b200", "_lib", "libhyperex_b200.so")


def disasm(fun):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout.splitlines()
    ins, i = [], 0
    while i < len(txt):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", txt[i])
        if m and i + 1 < len(txt):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", txt[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(),
                            dict(stall=(hi >> 41) & 0xf, wr=(hi >> 46) & 7, rd=(hi >> 49) & 7, wait=(hi >> 52) & 0x3f)))
                i += 2
                continue
        i += 1
    return ins


def main():
    fun = sys.argv[1]
    ins = disasm(fun)
    if not ins:
        raise SystemExit(f"{fun}: not found in {LIB}")
    addr = {a: i for i, (a, _, _) in enumerate(ins)}
    loops = []
    for i, (a, t, _) in enumerate(ins):
        m = re.search(r"BRA\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            n_lds = sum(1 for _, x, _ in ins[j:i + 1] if x.startswith("LDS.128") or "LDS.128" in x.split()[0:2])
            if n_lds:
                loops.append((i - j, j, i, n_lds))
    if not loops:
        raise SystemExit("no loop with LDS.128 found")
    _, j, i, n_lds = min(loops)
    if len(sys.argv) > 4:  # explicit address range (a kernel whose word loop is unrolled has no small loop)
        j, i = addr[int(sys.argv[3], 16)], addr[int(sys.argv[4], 16)]
        n_lds = sum(1 for _, x, _ in ins[j:i + 1] if "LDS.128" in x)
    body = ins[j:i + 1]
    whole = Counter((x.split()[1] if x.startswith("@") else x.split()[0]) for _, x, _ in ins)
    special = {k: v for k, v in whole.items() if k.startswith(("UBLKCP", "SYNCS", "LDG", "LDL", "STL", "S2R"))}
    ops = Counter((x.split()[1] if x.startswith("@") else x.split()[0]) for _, x, _ in body)
    out = [f"# {fun}", f"# hot loop 0x{ins[j][0]:x}..0x{ins[i][0]:x}: {len(body)} instructions, {n_lds} LDS.128 (Peq rows)",
           "# " + ", ".join(f"{k} {v}" for k, v in ops.most_common()),
           f"# whole kernel: {len(ins)} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(special.items())),
           "# addr   stall wr rd wait   instruction"]
    for a, t, c in body:
        out.append(f"0x{a:05x}  st{c['stall']:2d} w{c['wr'] if c['wr'] != 7 else '-'} r{c['rd'] if c['rd'] != 7 else '-'} {c['wait']:06b}  {t}")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
        print(out[1])
        print(out[2])
    else:
        print(text)


if __name__ == "__main__":
    main()
