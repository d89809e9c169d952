"""Join an ncu report's per-SASS-instruction metrics (--page source --csv) with nvdisasm line info of the built
library, and print instruction / stall-sample shares per source line and per function range.

usage: python tools/ncu_by_line.py gpurun_out/prof.ncu-rep 'k_integrateILi256E' [top_n]
"""
import csv, re, subprocess, sys, os, tempfile
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.environ.get("CT_ROOT", os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))  # CT_ROOT: tree whose built library was profiled
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s/chemtools_b200/libchemtools_b200.so > /dev/null" % (tmp, root), shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith("capi")][0]
sass = subprocess.check_output(["nvdisasm", "--print-line-info-inline" if os.environ.get("CT_INLINE") else "--print-line-info", os.path.join(tmp, cub)], text=True, stderr=subprocess.DEVNULL).split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l][0]
seq, cur, pending_reset = [], None, False
for l in sass[start + 1:]:
    if l.startswith(".text.") and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        # CT_INLINE=1: header intrinsics (shuffles, __syncwarp) are attributed to the kernels.cuh line they were inlined at
        if pending_reset:
            cur = None; pending_reset = False
        cand = (m.group(1).split("/")[-1], int(m.group(2)))
        if cur is None or (not cur[0].endswith("kernels.cuh") and cand[0].endswith("kernels.cuh")):
            cur = cand
        if not cur[0].endswith("kernels.cuh") and m.group(3) and m.group(3).endswith("kernels.cuh"):
            cur = (m.group(3).split("/")[-1], int(m.group(4)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
        pending_reset = True
csvtxt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(csvtxt.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
ci = {n: i for i, n in enumerate(hdr)}
base = int(data[0][0], 16)
bya = {s[0]: s for s in seq}
agg = defaultdict(lambda: [0, 0])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st_tot = defaultdict(int)
ops = defaultdict(lambda: [0, 0])
for r in data:
    s = bya.get(int(r[0], 16) - base)
    k = s[1] if s else None
    n, sm = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
    agg[k][0] += n; agg[k][1] += sm
    if os.environ.get("CT_STALL"):  # samples of one stall reason instead of all samples
        agg[k][1] += int(r[ci["stall_" + os.environ["CT_STALL"]]]) - sm
    for h in stalls:
        st_tot[h] += int(r[ci[h]])
    t = r[ci["Source"]].split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[o][0] += n; ops[o][1] += sm
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("kernel %s: %d SASS instructions, %.3g warp-instructions executed, %d samples" % (kern, len(seq), ti, ts))
T = sum(st_tot.values())
print("stall reasons:", ", ".join("%s %.1f%%" % (h.replace("stall_", ""), 100 * v / T) for h, v in sorted(st_tot.items(), key=lambda kv: -kv[1])[:8]))
print("opcodes:", ", ".join("%s %.1f%%/%.1f%%" % (o, 100 * v[0] / ti, 100 * v[1] / ts) for o, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:16]))
src = open(os.path.join(root, "chemtools_b200/csrc/kernels.cuh")).read().split("\n")
# function ranges from the source
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?__(?:device|global)__ .*?\b([A-Za-z_0-9:]+)\(", l)
    if m and not l.strip().startswith("//"):
        funcs.append((i, m.group(1)))
funcs.append((len(src) + 1, "end"))
print("\n-- by function (kernels.cuh)")
for (a, nm), (b, _) in zip(funcs, funcs[1:]):
    i = sum(v[0] for k, v in agg.items() if k and k[0] == "kernels.cuh" and a <= k[1] < b)
    s_ = sum(v[1] for k, v in agg.items() if k and k[0] == "kernels.cuh" and a <= k[1] < b)
    if s_ > 0.002 * ts:
        print("%-26s lines %4d-%4d  %5.1f%% inst %5.1f%% samples" % (nm, a, b - 1, 100 * i / ti, 100 * s_ / ts))
oth = defaultdict(lambda: [0, 0])
for k, v in agg.items():
    if not k or k[0] != "kernels.cuh":
        oth[k[0] if k else "none"][0] += v[0]; oth[k[0] if k else "none"][1] += v[1]
for f, v in oth.items():
    print("%-26s %5.1f%% inst %5.1f%% samples" % (f, 100 * v[0] / ti, 100 * v[1] / ts))
print("\n-- top lines by samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    line = src[k[1] - 1].strip()[:100] if k and k[0] == "kernels.cuh" and k[1] <= len(src) else ""
    print("%-28s %5.1f%% inst %5.1f%% samp  %s" % (k, 100 * v[0] / ti, 100 * v[1] / ts, line))
