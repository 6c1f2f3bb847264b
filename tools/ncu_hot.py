"""Top SASS instructions by stall samples from `ncu --page source --csv` output (reads the .ncu-rep here, no GPU).
   python tools/ncu_hot.py gpurun_out/x.ncu-rep <kernel index> [top N]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
view = sys.argv[4] if len(sys.argv) > 4 else "sass"
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view]
if kid != "all":
    cmd += ["--kernel-id", f":::{kid}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:150])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S = ix["# Samples"]
tot = sum(int(r[S]) for r in data if r[S].isdigit())
cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Global", "L2 Theoretical Sectors Global Ideal"]
print("total samples", tot, "instructions", sum(int(r[ix["Instructions Executed"]]) for r in data if r[ix["Instructions Executed"]].isdigit()))
stall_cols = [h for h in hdr if h.startswith("stall_")] 
rank = sorted(range(len(data)), key=lambda i: -int(data[i][S]) if data[i][S].isdigit() else 0)[:top]
for i in sorted(rank):
    r = data[i]
    extra = ""
    if stall_cols:
        st = sorted(((int(r[ix[c]]), c) for c in stall_cols if r[ix[c]].isdigit() and int(r[ix[c]]) > 0), reverse=True)[:2]
        extra = " ".join(f"{c[6:]}={v}" for v, c in st)
    print(f"{i:5d} {100.0*int(r[S])/tot:5.1f}%  {r[ix['Source']][:70]:70s} " + " ".join(f"{r[ix[c]]:>9s}" for c in cols[1:]) + "  " + extra)
