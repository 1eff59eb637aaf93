#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines: joins `ncu --page source --csv` (SASS view) with
`nvdisasm -g` line info of the in-tree library.  usage: ncu_lines.py report.ncu-rep kernel_substring [topN]"""
import csv, io, re, subprocess, sys, os, collections, glob, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if '# Samples' in r)
H = rows[hdr]
si, ai = H.index('# Samples'), H.index('Address')
stall_cols = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
samples = []
for r in rows[hdr + 1:]:
    try:
        samples.append((int(r[ai], 16) if r[ai].startswith('0x') else int(r[ai]), int(r[si] or 0), r))
    except Exception:
        pass
base = min(a for a, _, _ in samples)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "semiclassical_b200/lib/libsemiclassical_b200.so")], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(tmp + "/*.cubin")[0]], capture_output=True, text=True).stdout.splitlines()
# locate kernel section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.Counter()
stall = collections.defaultdict(collections.Counter)
tot = 0
for a, s, r in samples:
    key = line_of.get(a - base)
    agg[key] += s
    tot += s
    for c in stall_cols:
        try:
            v = int(r[c] or 0)
        except ValueError:
            v = 0
        if v:
            stall[key][H[c]] += v
src_cache = {}
def src(key):
    if key is None:
        return "?"
    f, ln = key
    if f not in src_cache:
        p = glob.glob(os.path.join(root, "semiclassical_b200/csrc", f))
        src_cache[f] = open(p[0]).read().splitlines() if p else []
    L = src_cache[f]
    return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
print(f"total samples {tot}")
for key, s in agg.most_common(top):
    st = ", ".join(f"{k[6:]}:{v}" for k, v in stall[key].most_common(3))
    print(f"{s:7d} {100.0*s/tot:5.1f}%  {key[0] if key else '?':16s}:{key[1] if key else 0:4d}  {src(key):90s} [{st}]")
