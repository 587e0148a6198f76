#!/usr/bin/env python
"""Turn one `ncu --set full` report into the markdown summary kept under profiles/.

    python tools/ncu_report_md.py gpurun_out/<tag>_full.ncu-rep profiles/<tag>_ncu_full.md [kernel-regex ...]

Per launch: duration, DRAM bytes read/written (the `traffic` of bench.py's roofline), SM / DRAM
throughput %, issue-slot utilisation, achieved occupancy, registers.  For every kernel regex given
(default: the two blend kernels) the SASS-level summary of tools/ncu_source_summary.py is appended.
"""
import csv
import io
import subprocess
import sys
import os

HERE = os.path.dirname(os.path.abspath(__file__))

COLS = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "ms"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("dram__bytes.sum.per_second", "dram GB/s"),
    ("__dram_pct_of_measured_peak", "% of 6549 GB/s"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    regexes = sys.argv[3:] or ["blend_forward", "blend_backward"]
    raw = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of `{os.path.basename(rep)}`", "",
             "Per launch (cold-cache, serialised by ncu: shares are meaningful, absolutes are not bench values).", "",
             "| " + " | ".join(n for _, n in COLS) + " |", "|" + "---|" * len(COLS)]
    for r in rows[2:]:
        cells = []
        for key, _ in COLS:
            v = r[ix[key]] if key in ix else ""
            if key == "__dram_pct_of_measured_peak" and "dram__bytes.sum.per_second" in ix:
                try:
                    bps = float(r[ix["dram__bytes.sum.per_second"]].replace(",", ""))
                    un = units[ix["dram__bytes.sum.per_second"]]
                    gbs = bps * {"byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}.get(un, 1.0)
                    v = f"{100.0 * gbs / 6549.0:.1f}"
                except ValueError:
                    v = ""
                cells.append(v)
                continue
            if key == "Kernel Name":
                v = v.split("(")[0].replace("brs::<unnamed>::", "").replace("void ", "")[:40]
            else:
                try:
                    f = float(v.replace(",", ""))
                    u = units[ix[key]]
                    if key.startswith("dram__bytes"):
                        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                        f *= scale
                    if key == "dram__bytes.sum.per_second":
                        scale = {"byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}.get(u, 1.0)
                        f *= scale
                    if key == "gpu__time_duration.sum":
                        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
                        f *= scale
                    v = f"{f:.4g}"
                except ValueError:
                    pass
            cells.append(v)
        lines.append("| " + " | ".join(cells) + " |")
    # secondary metrics the north star asks for (pipes, shared memory, L2 reductions), % of peak
    SEC = [("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe"),
           ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe"),
           ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU)"),
           ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU"),
           ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory data pipe"),
           ("lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "L2 reductions (RED)"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots")]
    picks = [r for r in rows[2:] if any(k in r[ix["Kernel Name"]] for k in ("blend_forward", "blend_backward", "preprocess_backward", "preprocess_kernel"))]
    if picks:
        lines += ["", "## Pipe / shared-memory / atomic utilisation (% of peak)", "",
                  "| kernel | " + " | ".join(n for _, n in SEC) + " |", "|" + "---|" * (len(SEC) + 1)]
        for r in picks:
            name = r[ix["Kernel Name"]].split("(")[0].replace("brs::<unnamed>::", "").replace("void ", "")[:40]
            vals = []
            for key, _ in SEC:
                try:
                    vals.append(f"{float(r[ix[key]].replace(',', '')):.1f}")
                except (KeyError, ValueError):
                    vals.append("")
            lines.append("| " + name + " | " + " | ".join(vals) + " |")
    for rx in regexes:
        src = run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"])
        if "Address" not in src:
            continue
        tmp = f"/tmp/_ncu_src_{rx}.csv"
        open(tmp, "w").write(src)
        lines += ["", f"## SASS-level profile: `{rx}`", "", "```",
                  run([sys.executable, os.path.join(HERE, "ncu_source_summary.py"), tmp, "25"]).rstrip(), "```"]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
