"""Top stall-sample SASS instructions of one kernel of an .ncu-rep (source page), to find the hot spots."""
import csv, subprocess, sys
rep, kid = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "::regex:gk_step:1")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("kernel", rows[0][1], "samples", tot, "warp-instructions", toti, "SASS lines", len(body))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100.0*int(r[ix['# Samples']])/tot:5.1f}% exec={int(r[ix['Instructions Executed']]):8d}  {r[ix['Source']].strip()[:70]:70s} {st}")
