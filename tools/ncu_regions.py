"""Sample share per block of SASS lines of one kernel, with marker instructions, from an .ncu-rep."""
import csv, subprocess, sys
rep, kid = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "0")
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", kid] if kid not in ("0", "") else []), capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
# keep only the first kernel instance if the regex matched several (addresses restart)
first = body[0][ix["Address"]]
for i in range(1, len(body)):
    if body[i][ix["Address"]] == first:
        body = body[:i]; break
tot = sum(int(r[ix["# Samples"]]) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("samples", tot, "lines", len(body))
marks = ("BAR.", "SYNCS", "UBLKCP", "EXIT", "ATOM", "MEMBAR", "RED.")
for b0 in range(0, len(body), blk):
    chunk = body[b0:b0 + blk]
    sm = sum(int(r[ix["# Samples"]]) for r in chunk)
    ex = max(int(r[ix["Instructions Executed"]] or 0) for r in chunk)
    st = {}
    for r in chunk:
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v: st[c[6:]] = st.get(c[6:], 0) + v
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    mk = sorted({m for r in chunk for m in marks if m in r[ix["Source"]]})
    nl = sum(1 for r in chunk if "LDS" in r[ix["Source"]]); ng = sum(1 for r in chunk if "LDG" in r[ix["Source"]]); nf = sum(1 for r in chunk if "DFMA" in r[ix["Source"]])
    if sm * 200 >= tot or mk:
        print(f"{b0:6d} {100.0*sm/tot:5.1f}% maxexec={ex:7d} LDS={nl:3d} LDG={ng:3d} DFMA={nf:3d} {top} {mk}")
