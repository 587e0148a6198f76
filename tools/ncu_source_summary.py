#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output of one kernel: executed instructions by opcode class and
the hottest SASS lines (by executed warp instructions and by stall samples).

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<name> > src.csv
    python tools/ncu_source_summary.py src.csv [top]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    # first row is the kernel name, second the header
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr) - 2]
    tot_inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
    tot_samp = sum(float(r[ix["# Samples"]] or 0) for r in data)
    print(rows[0][1] if len(rows[0]) > 1 else "", "| SASS lines", len(data), "| warp-instr executed %.3e | samples %d" % (tot_inst, tot_samp))
    by_op = collections.Counter()
    by_op_s = collections.Counter()
    for r in data:
        op = r[ix["Source"]].strip().split()
        op = [t for t in op if not t.startswith("@")]
        name = op[0].split(".")[0] if op else "?"
        by_op[name] += float(r[ix["Instructions Executed"]] or 0)
        by_op_s[name] += float(r[ix["# Samples"]] or 0)
    print("opcode        %instr  %samples")
    for k, v in by_op.most_common(22):
        print(f"  {k:10s} {100 * v / tot_inst:6.2f}  {100 * by_op_s[k] / max(tot_samp, 1):6.2f}")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_st = {h: sum(float(r[ix[h]] or 0) for r in data) for h in stall_cols}
    s = sum(tot_st.values())
    print("stalls:", ", ".join(f"{h[6:]} {100 * v / s:.1f}%" for h, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
    print(f"hottest {top} lines by samples:")
    for r in sorted(data, key=lambda r: -float(r[ix["# Samples"]] or 0))[:top]:
        st = sorted(((float(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"  {r[ix['Address']][-5:]} {100 * float(r[ix['# Samples']] or 0) / max(tot_samp, 1):5.2f}% inst={float(r[ix['Instructions Executed']] or 0):.2e} "
              f"thr={r[ix['Avg. Threads Executed']][:5]:>5s} {r[ix['Source']].strip()[:70]:70s} {st[0][1]}/{st[1][1]}")


if __name__ == "__main__":
    main()
