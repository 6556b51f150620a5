"""Print the hot SASS lines of an ncu source-page CSV (ncu -i x.ncu-rep --page source --csv)."""
import csv
import sys

path, nsym = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, ie, it, iw, isamp = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed",
                               "Avg. Threads Executed", "L1 Wavefronts Shared", "# Samples"))
tot = sum(int(r[ie]) for r in rows[2:] if r[ie].isdigit())
print("total warp instructions", tot, " per 1024 positions:", tot / nsym * 1024)
print("addr  exec/1024pos  avg_thr  samples  smem_wavefronts  sass")
for r in rows[2:]:
    if r[ie].isdigit() and int(r[ie]) > tot * thr:
        print(r[ia][-5:], f"{int(r[ie]) / nsym * 1024:7.2f}", f"{float(r[it]):5.1f}", r[isamp].rjust(6),
              r[iw].rjust(10), r[isrc][:100])
