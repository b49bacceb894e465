"""Per-SASS-instruction view of an ncu report: address, samples, executed count, top stall reasons.  Usage: ncu_sass.py rep kernel [start_hex end_hex]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi_ = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi_[0]]
iA, iS, iI, iSrc = h.index("Address"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [(i, n.replace("stall_", "")) for i, n in enumerate(h) if n.startswith("stall_")]
base = None
tot = 0
for r in rows[hi_[0] + 1:]:
    if len(r) <= iS or not r[iA].startswith("0x"):
        continue
    a = int(r[iA], 16)
    if base is None:
        base = a
    off = a - base
    try:
        s = int(r[iS] or 0)
    except ValueError:
        s = 0
    tot += s
    if lo <= off <= hi:
        st = []
        for i, n in stall_cols:
            try:
                v = int(r[i] or 0)
            except ValueError:
                v = 0
            if v:
                st.append((v, n))
        st.sort(reverse=True)
        print(f"{off:6x} {s:5d} {r[iI]:>9s}  {r[iSrc][:70]:70s} {' '.join(f'{n}:{v}' for v, n in st[:3])}")
print("total samples", tot)
