"""Attribute ncu SASS-level samples / executed instructions to source lines.

usage: ncu_lines.py <report.ncu-rep> <kernel mangled substring> [top]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the in-tree .so."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "midastouch_b200", "libmidas_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# collect (line info) per instruction of the kernel, in order
sections, cur, on = {}, None, None
for l in dis:
    if l.startswith("//--------------------- .text."):
        on = l if ksub in l else None
        if on:
            sections[on] = []
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), "inlined" in m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        sections[on].append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel instance only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[start]
body = [r for r in rows[start + 1:end] if r and r[0].startswith("0x")]
ci = {h: i for i, h in enumerate(hdr)}
cands = [v for v in sections.values() if len(v) == len(body)]
assert cands, (len(body), {k[:60]: len(v) for k, v in sections.items()})
lines = cands[0]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for r, li in zip(body, lines):
    s = int(r[ci["# Samples"]] or 0)
    ie = int(r[ci["Instructions Executed"]] or 0)
    agg[li[:2]][0] += s
    agg[li[:2]][1] += ie
    tot[0] += s
    tot[1] += ie
print("total samples", tot[0], "warp-instructions", tot[1])
src_cache = {}
def src(f, n):
    for d in ("midastouch_b200/csrc", "include"):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:90]
    return ""
for (f, n), (s, ie, _) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot[0]:5.1f}% samples {100*ie/tot[1]:5.1f}% inst  {f}:{n}  {src(f, n)}")
