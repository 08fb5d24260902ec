"""per-CUDA-source-line instruction counts / stall samples of one kernel of an .ncu-rep
   python tools/ncu_source.py rep.ncu-rep [kernel-index] [top-n]"""
import csv, subprocess, sys
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda"], capture_output=True, text=True).stdout
lines = out.splitlines()
# split into kernel sections
secs, cur = [], None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = {"name": ln, "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(ln)
sec = secs[kidx]
print(sec["name"][:150])
rows = list(csv.reader(sec["rows"]))
# find header of source table
hi = next(i for i, r in enumerate(rows) if r and r[0] in ("Line No", "#"))
h = rows[hi]
tot_inst = tot_samp = 0
data = []
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    d = dict(zip(h, r))
    try:
        inst = int(d.get("Instructions Executed", "0") or 0); samp = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    data.append((inst, samp, d.get("Line No", d.get("#")), d.get("Source", "")[:110]))
    tot_inst += inst; tot_samp += samp
print("total inst", tot_inst, "samples", tot_samp)
for inst, samp, ln, src in sorted(data, reverse=True)[:topn]:
    print(f"{100*inst/max(1,tot_inst):5.1f}% inst {100*samp/max(1,tot_samp):5.1f}% samp  L{ln:>4s} {src}")
