"""Summarise an ncu source-page CSV: executed warp-instructions by SASS opcode and by CUDA source line."""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
hi = his[0]
hdr = rows[hi]
iI, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
byop, stall, byline, stline = (collections.Counter() for _ in range(4))
tot = 0
cur_line = None
for r in rows[hi + 1:]:
    if len(r) <= iI:
        continue
    sass = r[3]
    try:
        n = int(r[iI])
    except ValueError:
        continue
    if sass == "" or sass == "-":       # a CUDA source line row: remember it
        continue
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    byop[op] += n
    tot += n
    try:
        stall[op] += int(r[iS])
    except ValueError:
        pass
    byline[r[1][:70]] += n
print("total warp-instructions", tot)
for k, v in byop.most_common(28):
    print(f"  {k:10s} {v:12d} {100 * v / tot:5.1f}%  stall {stall[k]}")
