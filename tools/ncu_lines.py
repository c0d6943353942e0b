#!/usr/bin/env python
"""tools/ncu_lines.py REPORT KERNEL_REGEX [TOP] -- per-CUDA-source-line totals (stall samples, warp
instructions executed) from an `ncu --set full --import-source on` report, read without a GPU via
`ncu -i ... --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, agg, tot_s, tot_i = "", [], 0, 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        if len(r) > 8 and r[0].isdigit() and r[4].isdigit():
            s, i = int(r[6]), int(r[7])
            agg.append((s, i, cur_file, int(r[0]), r[1].strip()[:100]))
            tot_s += s; tot_i += i
    agg.sort(reverse=True)
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for s, i, f, ln, src in agg[:top]:
        print(f"{100.0 * s / max(tot_s, 1):5.1f}% smp {100.0 * i / max(tot_i, 1):5.1f}% inst  {f}:{ln}  {src}")


if __name__ == "__main__":
    main()
