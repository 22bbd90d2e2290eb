"""Top stall-sample SASS instructions of an .ncu-rep source page: python profiles/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# a report with several launches repeats the header: keep the first launch's table only (or the one named by argv[3])
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
start = starts[which]
end = next((i for i, l in enumerate(lines[start + 1:], start + 1) if l.startswith('"Kernel Name"')), len(lines))
r = list(csv.DictReader(lines[start:end]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(x["# Samples"] or 0) for x in r)
print("total samples", tot, "instructions", len(r))
stall_cols = [c for c in r[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
idx = {id(x): i for i, x in enumerate(r)}
for x in sorted(r, key=lambda x: -int(x["# Samples"] or 0))[:n]:
    s = int(x["# Samples"] or 0)
    top = sorted(((int(x[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{idx[id(x)]:5d} {100.0*s/tot:5.1f}% exec={x['Instructions Executed']:>8s} {x['Source'][:90]:90s} {top}")
