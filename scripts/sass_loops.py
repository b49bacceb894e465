"""Static SASS summary of one kernel: total opcode histogram and, for every backward branch (loop), the opcode histogram
of the loop body.  Usage: python scripts/sass_loops.py <obj-or-so> <mangled-name-substring> [min_body_len]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []   # (addr, text)
    for line in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    def op(t):
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        return t.split()[0].split(".")[0]
    print("==", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:140], "instructions:", len(ins))
    addr2idx = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        if op(t) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) in addr2idx and int(m.group(1), 16) <= a:
                j = addr2idx[int(m.group(1), 16)]
                body = ins[j:i + 1]
                if len(body) < minlen:
                    continue
                h = collections.Counter(op(x) for _, x in body)
                print(f"  loop {ins[j][0]:#x}..{a:#x}: {len(body)} instr:", ", ".join(f"{k} {v}" for k, v in h.most_common(24)))
    break
