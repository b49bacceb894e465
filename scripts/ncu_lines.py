"""Top CUDA source lines by stall samples / executed instructions from an ncu report (SASS rows grouped by line)."""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
smp, ins, txt = collections.Counter(), collections.Counter(), {}
cur = "?"
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    elif len(r) > 8 and r[0].isdigit() and r[2] == "-":
        key = (cur, int(r[0]))
        txt[key] = r[1].strip()[:95]
        try:
            smp[key] += int(r[iS] or 0); ins[key] += int(r[iI] or 0)
        except ValueError:
            pass
ts, ti = sum(smp.values()) or 1, sum(ins.values()) or 1
print("total samples", ts, "total warp-instr", ti)
for key, s in smp.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    print(f"{100*s/ts:5.1f}% smp {100*ins[key]/ti:5.1f}% ins  {key[0]}:{key[1]}  {txt[key]}")
