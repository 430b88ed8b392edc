"""Per-instruction executed counts from an ncu report's source page: prints regions of the SASS with their
share of executed warp instructions (to find where the issue slots go)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ie = hdr.index("Instructions Executed"); src = hdr.index("Source"); smp = hdr.index("# Samples")
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name": break
    data.append((r[src].strip(), int(r[ie]), int(r[smp])))
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print("instructions", len(data), "executed", tot, "samples", tots)
mode = sys.argv[2] if len(sys.argv) > 2 else "blocks"
if mode == "all":
    for i, (s, n, k) in enumerate(data):
        print(f"{i:5d} {n:8d} {k:5d}  {s}")
else:
    # group consecutive instructions with equal execution count
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and data[j + 1][1] == data[i][1]: j += 1
        n = data[i][1] * (j - i + 1)
        sm = sum(d[2] for d in data[i:j + 1])
        if n > tot * 0.004:
            ops = {}
            for s, _, _ in data[i:j + 1]:
                op = s.split()[0] if not s.startswith("@") else s.split()[1]
                op = op.split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:8]
            print(f"[{i:5d}-{j:5d}] len {j-i+1:4d} x {data[i][1]:7d} = {n/tot:6.1%} instr, {sm/max(tots,1):6.1%} samples  {top}")
        i = j + 1
