"""Static size of the HOT code: SASS instructions of one kernel that were executed at least `thr` times (from an ncu
report's source page), grouped by source function.  usage: hot_footprint.py report.ncu-rep kernel [thr]"""
import csv, re, subprocess, sys, os, tempfile
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e6
root = os.environ.get("CT_ROOT", os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s/chemtools_b200/libchemtools_b200.so > /dev/null" % (tmp, root), shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith("capi")][0]
sass = subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], text=True, stderr=subprocess.DEVNULL).split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l][0]
seq, cur, sec = [], None, "main"
for l in sass[start + 1:]:
    if l.startswith(".text.") and seq: break
    m = re.match(r"\s*(\$[^:]+):", l)
    if m: sec = m.group(1).split("$")[-1][:40]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((int(m.group(1), 16), cur, sec))
rows = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL).split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 10]
ci = {n: i for i, n in enumerate(hdr)}
base = int(data[0][0], 16)
ex = {int(r[0], 16) - base: int(r[ci["Instructions Executed"]]) for r in data}
src = open(os.path.join(root, "chemtools_b200/csrc/kernels.cuh")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?__(?:device|global)__ .*?\b([A-Za-z_0-9:]+)\(", l)
    if m and not l.strip().startswith("//"): funcs.append((i, m.group(1)))
def fn(k):
    if not k or not k[0].endswith("kernels.cuh"): return k[0] if k else "?"
    name = "?"
    for a, n in funcs:
        if a <= k[1]: name = n
        else: break
    return name
hot = defaultdict(int); tot = defaultdict(int); bysec = defaultdict(lambda: [0, 0])
lines_hot = set()
for a, k, sec in seq:
    f = fn(k); tot[f] += 1; bysec[sec][1] += 1
    if ex.get(a, 0) >= thr:
        hot[f] += 1; bysec[sec][0] += 1; lines_hot.add(a // 128)
print("hot instructions (executed >= %.0e): %d of %d; distinct 128-byte lines touched: %d (%.0f KB)" % (thr, sum(hot.values()), len(seq), len(lines_hot), len(lines_hot) * 128 / 1024))
print("-- by section (hot / all)")
for k, v in sorted(bysec.items(), key=lambda kv: -kv[1][0])[:14]: print("%6d / %6d  %s" % (v[0], v[1], k))
print("-- by function (hot / all)")
for k, v in sorted(hot.items(), key=lambda kv: -kv[1])[:30]: print("%6d / %6d  %s" % (v, tot[k], k))
