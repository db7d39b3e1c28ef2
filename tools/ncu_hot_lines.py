"""Executed warp-instructions per CUDA source line of one kernel: joins the SASS page of an .ncu-rep (per-instruction
counts) with nvdisasm's line table of the object the kernel was built from (same instruction order).
python tools/ncu_hot_lines.py rep kernel_regex object.o [top_n]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile

rep, kern, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
kname = rows[hi - 1][1] if hi > 0 else kern
sass = [r for r in rows[hi + 1:] if len(r) > 6 and r[0].startswith("0x")]
if len([i for i, r in enumerate(rows) if r and r[0] == "Address"]) > 1:      # first kernel only
    nxt = [i for i, r in enumerate(rows) if r and r[0] == "Address"][1]
    sass = [r for r in rows[hi + 1:nxt] if len(r) > 6 and r[0].startswith("0x")]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
tmp = tempfile.mkdtemp()
import os
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
import glob
cubin = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function whose mangled name matches
# mangled-name fragments of the kernel: base name, and for a template instance its Itanium argument list (LiNE / LbNE)
plain = kname.split("(const")[0] if "<" in kname else re.sub(r"\(.*", "", kname)
fn_short = re.sub(r"<.*", "", re.sub(r"^void ", "", plain)).split("::")[-1]
targs = ""
m = re.search(r"<(.*)>", plain)
if m:
    parts = re.findall(r"\((int|bool)\)\s*(-?\d+)", m.group(1))
    targs = "I" + "".join(("Lb%sE" if kind == "bool" else "Li%sE") % val for kind, val in parts) + "E"
lines_of = []
cur_line, in_fn = None, False
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        in_fn = fn_short in m.group(1) and targs in m.group(1)
        continue
    if ln.startswith("\t.section") or ln.startswith(".section"):
        in_fn = False
    if not in_fn:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1), int(m.group(2)))      # inlined helpers live in other files (psg_common.cuh ...): keep the file
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_of.append(cur_line)
print(f"kernel {kname[:80]}: {len(sass)} SASS rows in the report, {len(lines_of)} instructions in the object")
per, smp = collections.Counter(), collections.Counter()
total = 0
for i, r in enumerate(sass):
    n = int(r[ie])
    line = lines_of[i] if i < len(lines_of) and lines_of[i] is not None else -1
    per[line] += n
    smp[line] += int(r[isamp] or 0)
    total += n
sources = {}


def text_of(key):
    if not isinstance(key, tuple):
        return ""
    path, line = key
    if path not in sources:
        try:
            sources[path] = open(path).read().splitlines()
        except OSError:
            sources[path] = []
    lines = sources[path]
    return lines[line - 1].strip() if 0 < line <= len(lines) else ""


print(f"total warp-instructions {total}")
order = sorted(per, key=lambda l: -smp[l]) if "--by-samples" in sys.argv else [l for l, _ in per.most_common()]
for key in order[:top]:
    where = f"{key[0].split('/')[-1]}:{key[1]}" if isinstance(key, tuple) else "?"
    print(f"{per[key] / total * 100:5.1f}%  samples {smp[key]:6d}  {where}: {text_of(key)[:110]}")
