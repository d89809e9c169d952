"""Static SASS listing / statistics of one out-of-line device function inside one kernel of the built library.
usage: python tools/sass_fn.py <kernel substring> <function substring> [list]"""
import re, subprocess, sys, os, tempfile, collections
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
kern, fn = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.check_call("cd %s && cuobjdump -xelf all %s/chemtools_b200/libchemtools_b200.so > /dev/null" % (tmp, root), shell=True)
cub = [f for f in os.listdir(tmp) if f.startswith("capi")][0]
sass = subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], text=True, stderr=subprocess.DEVNULL).split("\n")
inside = False; cur = None; out = []; cnt = collections.Counter(); ops = collections.Counter()
for l in sass:
    if l.startswith(".text.") or (l.startswith("$") and l.rstrip().endswith(":")):
        inside = (kern in l) and (fn in l) and l.startswith("$")
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        cnt[cur] += 1; ops[m.group(1).split(".")[0]] += 1
        out.append("%-28s %s" % ("%s:%d" % cur if cur else "", l.strip()[:120]))
    elif re.match(r"\.L_x", l): out.append(l.strip())
print("total", sum(cnt.values()))
print(ops.most_common(30))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:25]: print(v, k)
if len(sys.argv) > 3: print("\n".join(out))
