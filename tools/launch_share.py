#!/usr/bin/env python
"""Compare the ncu launch list of the bench command with the stage times bench.py measured live.

    python tools/launch_share.py gpurun_out/<tag>_launches_bench.csv gpurun_out/<tag>_bench_ours.json profiles/<tag>_launch_share.md

ncu's per-launch times are cold-cache and serialised, so only each kernel's SHARE of the native-kernel
time is compared with the share of the CUDA-event stage times (roofline.stages of the bench line)."""
import collections
import csv
import json
import os
import sys

STAGE_OF = {
    "preprocess_kernel": "preprocess", "preprocess_backward_kernel": "preprocess_bwd", "blend_forward_kernel": "blend_fwd",
    "blend_backward_kernel": "blend_bwd", "emit_kernel": "coarse_emit", "fine_kernel": "fine_bin", "fine_scan_kernel": "fine_bin",
    "fine_plan_kernel": "fine_bin", "tile_offsets_kernel": "fine_bin", "tile_ranges_kernel": "fine_bin",
    "upsweep_kernel": "sorts", "scan_kernel": "sorts", "downsweep_kernel": "sorts",
    "onesweep_kernel": "sorts", "radix_hist_kernel": "sorts",
}


def main():
    launches, bench, out = sys.argv[1:4]
    rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t = collections.defaultdict(float)
    n = collections.Counter()
    for r in rows[1:]:
        name = r[ki].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        t[name] += float(r[vi]) * 1e-3
        n[name] += 1
    total = sum(t.values())
    d = json.loads(open(bench).read().strip().splitlines()[-1])
    stages = {k: v["ms"] for k, v in d["roofline"]["stages"].items()}
    stages["sorts"] = stages.pop("depth_sort") + stages.pop("coarse_sort")
    ssum = sum(stages.values())
    by_stage = collections.defaultdict(float)
    for k, v in t.items():
        by_stage[STAGE_OF.get(k, k)] += v
    lines = [f"# Launch list of `python bench.py --steps 1 --warmup 3` under ncu vs the live stage times", "",
             f"{sum(n.values())} native-kernel launches captured in steady state (`--launch-skip 1200 --launch-count 400`; file "
             f"`profiles/{os.path.basename(out).replace('_launch_share.md', '_launches_bench.csv')}`), "
             f"`ncu --metrics gpu__time_duration.sum --clock-control none`.  Per-launch times are cold-cache and serialised "
             "by ncu, so only the SHARE of each kernel is comparable with the CUDA-event stage times that `bench.py` "
             f"measures live (`roofline.stages`, bench line of the same GPU job: {d['value']:.1f} views/s).", "",
             "| kernel | launches | avg µs | share of native-kernel time (ncu) |", "|---|---|---|---|"]
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        lines.append(f"| `{k}` | {n[k]} | {v / n[k]:.1f} | {100 * v / total:.1f} % |")
    lines += ["", "| stage | share under ncu | share of the stage-time sum (bench, CUDA events) | stage ms (bench) |", "|---|---|---|---|"]
    for k in sorted(stages, key=lambda k: -stages[k]):
        lines.append(f"| {k} | {100 * by_stage.get(k, 0.0) / total:.1f} % | {100 * stages[k] / ssum:.1f} % | {stages[k]:.4f} |")
    dom = d["roofline"]["kernel"]
    lines += ["", f"Dominant kernel `{dom}`: {100 * by_stage.get(dom, 0.0) / total:.1f} % of native-kernel time under ncu, "
              f"{100 * stages[dom] / ssum:.1f} % of the stage-time sum in the bench run "
              "(`sorts` = depth sort + coarse sort passes, which share the radix kernels)."]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
