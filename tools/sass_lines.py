"""Static SASS instruction count per source line inside one section (noinline function) of a kernel.
usage: python tools/sass_lines.py k_integrateILi384E rhs_entryE [top_n]"""
import re, subprocess, sys, os, tempfile
from collections import defaultdict
kern, sect = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.environ.get("CT_ROOT", os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s/chemtools_b200/libchemtools_b200.so > /dev/null" % (tmp, root), shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith("capi")][0]
sass = subprocess.check_output(["nvdisasm", "--print-line-info-inline" if os.environ.get("CT_INLINE") else "--print-line-info", os.path.join(tmp, cub)], text=True, stderr=subprocess.DEVNULL).split("\n")
src = open(os.path.join(root, "chemtools_b200/csrc/kernels.cuh")).read().split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l][0]
cnt = defaultdict(int); ops = defaultdict(lambda: defaultdict(int)); cur = None; insec = (sect == "main"); tot = 0; pending = False
for l in sass[start + 1:]:
    if l.startswith(".text.") and tot: break
    m = re.match(r"\s*(\$[^:]+):", l)
    if m:
        insec = sect in m.group(1) if sect != "main" else False
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if pending: cur = None; pending = False
        cand = (os.path.basename(m.group(1)), int(m.group(2)))
        if cur is None or (not cur[0].endswith("kernels.cuh") and cand[0].endswith("kernels.cuh")): cur = cand
        if not cur[0].endswith("kernels.cuh") and m.group(3) and m.group(3).endswith("kernels.cuh"): cur = (os.path.basename(m.group(3)), int(m.group(4)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", l)
    if m:
        pending = True
        if insec:
            cnt[cur] += 1; tot += 1; ops[cur][m.group(1).split(".")[0].lstrip("@!P0123456789 ")] += 1
print("section", sect, "instructions", tot)
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:topn]:
    line = src[k[1] - 1].strip()[:100] if k and k[0].endswith("kernels.cuh") and k[1] <= len(src) else ""
    print("%5d  %s:%d  %s" % (v, k[0] if k else "?", k[1] if k else 0, line))
