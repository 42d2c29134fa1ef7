"""Turn an .ncu-rep (ncu --set full) into the small text/JSON summaries committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_v0.ncu-rep profiles/r01 [--traffic-kernel joint_hist]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_global_red.sum", "sm__cycles_active.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    tk = sys.argv[sys.argv.index("--traffic-kernel") + 1] if "--traffic-kernel" in sys.argv else None
    hdr, units, rows = raw(rep)
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none --import-source on  ({Path(rep).name})",
             "# one column per captured launch; values are per launch", ""]
    names = [r[ki].split("(")[0].replace("void ", "").replace("nmi::<unnamed>::", "") for r in rows]
    lines.append("kernel".ljust(66) + " | " + " | ".join(n[:34].ljust(34) for n in names))
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            lines.append(f"{m} [{units[i]}]".ljust(66) + " | " + " | ".join(r[i][:34].ljust(34) for r in rows))
    Path(prefix + "_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    if tk:
        for r in rows:
            if tk in r[ki]:
                def val(m):
                    i = hdr.index(m)
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
                    return float(r[i]) * scale
                d = {"kernel": r[ki].split("(")[0], "dram_bytes_read": val("dram__bytes_read.sum"),
                     "dram_bytes_write": val("dram__bytes_write.sum"), "source": Path(rep).name}
                d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]
                Path(prefix).parent.joinpath("hist_kernel_traffic.json").write_text(json.dumps(d, indent=1))
                break
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    keep = [l for l in det.splitlines() if not l.strip().startswith(("OPT", "INF", "http"))]
    Path(prefix + "_ncu_details.txt").write_text("\n".join(keep[:1200]) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
