"""SASS listing with executions per unit:  python tools/ncu_sass.py rep kernel_index units [min_per_unit]
prints runs of consecutive instructions with the same execution count: [count/unit] n_instr  first..last opcode"""
import csv, subprocess, sys
rep, kidx, units = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
secs, cur = [], None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = []
        secs.append((ln, cur))
    elif cur is not None:
        cur.append(ln)
name, body = secs[kidx]
print(name[:120])
rows = list(csv.reader(body))
h = rows[0]
ii, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
runs = []
tot = 0; tots = 0
for r in rows[1:]:
    if len(r) <= ii: continue
    c = int(r[ii]); s = int(r[ss]); op = r[si].split()[0] if r[si].split() else "?"
    if op.startswith("@"): op = r[si].split()[1]
    tot += c; tots += s
    if runs and runs[-1][0] == c:
        runs[-1][1] += 1; runs[-1][3] = op; runs[-1][4] += s; runs[-1][5].append(op)
    else:
        runs.append([c, 1, op, op, s, [op]])
print("total instr/unit", tot / units, "samples", tots)
for c, n, a, b, s, ops in runs:
    if c / units * n >= thr:
        import collections
        cnt = collections.Counter(o.split(".")[0] for o in ops)
        print(f"{c/units:8.2f}/unit x {n:3d} instr = {c/units*n:8.1f}  samp {100*s/max(1,tots):5.1f}%  " + " ".join(f"{k}:{v}" for k, v in cnt.most_common(8)))
