"""Turns the .ncu-rep files brought back in gpurun_out/ into the small CSV/JSON summaries committed under
profiles/ (run here, no GPU needed):  python tools/ncu_summary.py <round-tag> [bases_per_launch]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_alu",
        "sm__pipe_alu_cycles_active", "sm__pipe_fma", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu",
        "smsp__issue_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed", "launch__",
        "smsp__average_warps_issue_stalled", "sm__warps_active", "sm__throughput", "gpu__dram_throughput",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld", "l1tex__t_sectors_pipe_lsu_mem_local",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sass__inst_executed_local", "sass__inst_executed_shared",
        "sm__cycles_elapsed", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    bases = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for name, rep in (("myers_pair_kernel_u32_15", "prof_myers"), ("classify_kernel", "prof_classify")):
        path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        hdr, units, vals = raw(path)
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.csv"), "w") as f:
            f.write("metric,value,unit\n")
            for h, u, v in zip(hdr, units, vals):
                if any(h.startswith(k) for k in KEEP) and "per_second" not in h:
                    f.write(f"{h},{v},{u}\n")
        if rep == "prof_myers":
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

            def b(key):
                v, u = d[key]
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

            json.dump({"kernel": "hx_myers_pair_kernel<unsigned int, 15, false>", "bases_per_launch": bases,
                       "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                       "gpu_time_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")) *
                       {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[d["gpu__time_duration.sum"][1]],
                       "alu_pipe_pct_of_peak_active": float(
                           d["sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"][0]),
                       "source": f"ncu --set full --clock-control none, one launch of `python bench.py --steps 1 "
                                 f"--warmup 1 --no-cpu-baseline --no-ab` ({tag})"},
                      open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
