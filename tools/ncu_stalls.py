"""Bucketed warp-stall summary of an `ncu --set full --import-source on` report (one kernel):
   python tools/ncu_stalls.py report.ncu-rep [bucket]   -> % of stall samples per SASS bucket + dominant reasons."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
si = hdr.index("Warp Stall Sampling (All Samples)")
reason_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = [r[1].strip() for r in data]
tot = sum(int(r[si]) for r in data if r[si].isdigit())
print(rows[0][1][:120])
for key in ("LDTM", "STTM", "MUFU.EX2", "UTCHMMA", "UTMALDG", "SYNCS.PHASECHK", "BAR."):
    ix = [i for i, o in enumerate(ops) if key in o]
    if ix:
        print(f"  {key:16s} first {ix[0]:5d} last {ix[-1]:5d} count {len(ix)}")
print(f"total samples {tot}, {len(ops)} SASS instructions, bucket {B}")
for b in range(0, len(data), B):
    chunk = data[b:b + B]
    sm = sum(int(r[si]) for r in chunk if r[si].isdigit())
    if sm < tot * 0.01:
        continue
    agg = collections.Counter()
    for r in chunk:
        for i in reason_cols:
            if r[i].isdigit():
                agg[hdr[i].replace("stall_", "")] += int(r[i])
    kinds = collections.Counter((o.split()[1] if o.startswith("@") and len(o.split()) > 1 else o.split()[0]).split(".")[0]
                                for o in (r[1].strip() for r in chunk) if o)
    print(f"[{b:4d}-{b + B:4d}] {100 * sm / tot:5.1f}%  {agg.most_common(3)}  {kinds.most_common(4)}")
