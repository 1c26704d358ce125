"""Turns the ncu exports brought back in gpurun_out/ (prof_<tag>.raw.csv from `ncu --page raw --csv`, written by
tools/profile_r02.sh / profile_fast.sh) into the small summaries committed under profiles/ (run here, no GPU needed):
    python tools/ncu_summary.py <round-tag> [bases_per_launch_of_the_1M_capture]
For every export: profiles/<tag>_ncu_<name>.csv (the metrics that decide what bounds the kernel) and, from the source page
when present, profiles/<tag>_hot_<name>.txt (the SASS instructions grouped by how often they executed).  The 1 M-read
capture of the graded kernel also gives profiles/<tag>_ncu_traffic.json (dram bytes per launch, bench.py's `traffic`)."""
import csv
import glob
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
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "sass__inst_executed_local",
        "sass__inst_executed_shared", "sm__cycles_elapsed", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    bases = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.raw.csv"))):
        name = os.path.basename(path)[len("prof_"):-len(".raw.csv")]
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.csv"), "w") as f:
            f.write(f"# {d.get('Kernel Name', ('?',))[0]}\n")
            f.write("metric,value,unit\n")
            for h, u, v in zip(hdr, units, vals):
                if any(h.startswith(k) for k in KEEP) and "per_second" not in h:
                    f.write(f"{h},{v},{u}\n")
        src = path[:-len(".raw.csv")] + ".src.csv"
        if os.path.exists(src):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), src, "10"], capture_output=True,
                                 text=True).stdout
            open(os.path.join(ROOT, "profiles", f"{tag}_hot_{name}.txt"), "w").write(
                f"# {d.get('Kernel Name', ('?',))[0]}\n# SASS instructions grouped by how often they executed (ncu source page)\n" + out)
        if name == "myers_1M":

            def b(key):
                v, u = d[key]
                return float(v.replace(",", "")) * UNIT[u]

            json.dump({"kernel": d["Kernel Name"][0], "bases_per_launch": bases,
                       "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                       "gpu_time_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")) * TIME[d["gpu__time_duration.sum"][1]],
                       "alu_pipe_pct_of_peak_active": float(d["sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"][0]),
                       "source": f"ncu --set full --clock-control none, one launch of `python bench.py --steps 2 --warmup 1 "
                                 f"--no-cpu-baseline --no-ab --no-file-e2e` ({tag})"},
                      open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
