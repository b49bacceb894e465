"""Top stalled SASS instructions of one kernel of an ncu report.  Usage: ncu_top.py rep launch_id [N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", kid, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[hi - 1][:2])
h = rows[hi]
iS, iI, iSrc = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Source")
stall = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") or n.lower().startswith("stall")]
extra = [(i, n) for i, n in enumerate(h) if n in ("L1 Tag Requests Global", "L2 Theoretical Sectors Global", "L1 Wavefronts Shared")]
recs = []
tot = 0
base = None
for r in rows[hi + 1:]:
    if len(r) <= iS or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    s = int(r[iS] or 0)
    tot += s
    st = sorted(((int(r[i] or 0), n) for i, n in stall if (r[i] or "0").isdigit() and int(r[i] or 0)), reverse=True)
    ex = " ".join(f"{n.split()[1] if ' ' in n else n}:{r[i]}" for i, n in extra if r[i] not in ("0", "", "-"))
    recs.append((s, a - base, r[iI], r[iSrc].strip()[:60], " ".join(f"{n.replace('stall_', '')}:{v}" for v, n in st[:3]), ex))
print("total samples", tot, "instructions", len(recs))
for s, off, ie, src, st, ex in sorted(recs, reverse=True)[:N]:
    print(f"{off:6x} {s:6d} {100.0 * s / max(tot, 1):5.1f}% exec {ie:>8s}  {src:60s} {st}  {ex}")
# totals of the request columns
for i, n in extra:
    t = 0
    for r in rows[hi + 1:]:
        try:
            t += int(r[i])
        except (ValueError, IndexError):
            pass
    print(n, t)
