#!/usr/bin/env python
"""Per-instruction view of a kernel in an .ncu-rep (captured with --import-source on), folded into regions of equal execution count:
share of the issued warp instructions, lanes per instruction, lost lane slots, stall samples.  Read here, no GPU needed.
  python tools/ncu_source_regions.py gpurun_out/x.ncu-rep [kernel index = 0] [min share % = 0.3]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    blk = rows[starts[which]:starts[which + 1]] if which + 1 < len(starts) else rows[starts[which]:]
    hdr, data = blk[1], blk[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    inst = lambda r: int(r[ix["Instructions Executed"]])
    thr = lambda r: int(r[ix["Thread Instructions Executed"]])
    smp = lambda r: int(r[ix["# Samples"]])
    tot_i, tot_t, tot_s = sum(map(inst, data)), sum(map(thr, data)), max(sum(map(smp, data)), 1)
    print(f"## {blk[0][1]}\n")
    print(f"{len(data)} SASS lines, {tot_i} warp instructions executed, {tot_t / tot_i:.2f} lanes per instruction\n")
    print("| SASS lines | instructions | share of issued | executions per line | lanes / instruction | lane slots lost (of all) | stall samples | first instruction |")
    print("|---|---|---|---|---|---|---|---|")
    prev, start, ai, at, asmp = None, 0, 0, 0, 0

    def flush(s, e):
        if ai < tot_i * min_share / 100:
            return
        print(f"| {s}-{e} | {e - s + 1} | {100 * ai / tot_i:.1f} % | {ai / (e - s + 1) / 1e6:.2f} M | {at / ai:.1f} | "
              f"{100 * (32 * ai - at) / (32 * tot_i):.2f} % | {100 * asmp / tot_s:.1f} % | `{data[s][ix['Source']].strip()[:48]}` |")

    for i, r in enumerate(data):
        ie = inst(r)
        if prev is not None and (ie > prev * 1.2 or ie < prev * 0.83):
            flush(start, i - 1)
            start, ai, at, asmp = i, 0, 0, 0
        ai += ie; at += thr(r); asmp += smp(r); prev = max(ie, 1)
    flush(start, len(data) - 1)


if __name__ == "__main__":
    main()
