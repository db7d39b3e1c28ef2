"""Stall-reason totals and the top stalled SASS instructions of one kernel in an .ncu-rep (source page).
python tools/ncu_stalls.py rep kernel_regex [top_n]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[0]
end = heads[1] if len(heads) > 1 else len(rows)
hdr = rows[hi]
sass = [r for r in rows[hi + 1:end] if len(r) > 6 and r[0].startswith("0x")]
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
isamp, iexec = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = collections.Counter()
for r in sass:
    for i, h in stall_cols:
        tot[h] += int(r[i] or 0)
all_s = sum(tot.values())
print(f"{kern}: {len(sass)} instructions, {all_s} samples, {sum(int(r[iexec]) for r in sass)} warp-instructions")
for h, n in tot.most_common(10):
    print(f"  {h:28s} {n:8d} {100.0 * n / max(all_s, 1):5.1f}%")
print("top instructions by samples:")
for r in sorted(sass, key=lambda r: -int(r[isamp] or 0))[:top]:
    reasons = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print(f"  {int(r[isamp]):6d}  exec {int(r[iexec]):9d}  {r[1][:70]:70s} {reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}")
