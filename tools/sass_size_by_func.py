"""Static SASS instruction count per source function (from nvdisasm line info) for one kernel."""
import re, subprocess, sys, os, tempfile
from collections import defaultdict
kern = sys.argv[1]
root = os.environ.get("CT_ROOT", os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))  # CT_ROOT: tree whose built library was profiled
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s/chemtools_b200/libchemtools_b200.so > /dev/null" % (tmp, root), shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith("capi")][0]
sass = subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], text=True, stderr=subprocess.DEVNULL).split("\n")
src = open(os.path.join(root, "chemtools_b200/csrc/kernels.cuh")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?__(?:device|global)__ .*?\b([A-Za-z_0-9:]+)\(", l)
    if m and not l.strip().startswith("//"):
        funcs.append((i, m.group(1)))
def fn(line):
    name = "?"
    for a, n in funcs:
        if a <= line: name = n
        else: break
    return name
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l][0]
cnt = defaultdict(int); cur = None; tot = 0; section = "main"
secs = defaultdict(int)
for l in sass[start + 1:]:
    if l.startswith(".text.") and tot: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = fn(int(m.group(2))) if m.group(1).endswith("kernels.cuh") else os.path.basename(m.group(1)); continue
    m = re.match(r"\s*(\$[^:]+|\.L_x_\d+):", l)
    if m and m.group(1).startswith("$"): section = m.group(1)[:80]
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l):
        cnt[cur] += 1; tot += 1; secs[section] += 1
print("total", tot)
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]: print("%6d %5.1f%%  %s" % (v, 100 * v / tot, k))
print("-- sections"); 
for k, v in sorted(secs.items(), key=lambda kv: -kv[1])[:12]: print("%6d  %s" % (v, k))
