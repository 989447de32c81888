"""Instructions executed and stall samples per device function: joins the ncu SASS page of one kernel with the line table
of its cubin (nvdisasm -g) and the function definitions of the CUDA source.
usage: func_profile.py <report.ncu-rep> <lib.so> <kernel-substring> <source.cuh> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kname, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
dis = ""
for f in sorted(os.listdir(d)):
    if f.endswith(".cubin"):
        t = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode(errors="replace")
        if kname in t:
            dis = t
            break
base_name = os.path.basename(srcfile)
funcs = []
for i, l in enumerate(open(srcfile), 1):
    m = re.match(r"^(?:__device__|__global__|static|__host__|template).*?\b([A-Za-z_]\w+)\s*\(", l)
    if m and not l.strip().endswith(";") and m.group(1) not in ("__launch_bounds__",):
        funcs.append((i, m.group(1)))
    m = re.match(r"^__global__.*\b(\w+)\(const", l)
    if m: funcs.append((i, m.group(1)))
funcs.sort()
def fn(line):
    name = "?"
    for s, n in funcs:
        if s <= line: name = n
        else: break
    return name
line_of, cur, infunc = {}, None, False
for l in dis.split("\n"):
    if l.startswith(".text.") or (".section" in l and ".text." in l): infunc = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and infunc: line_of[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode(errors="replace")
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] if len(his) > 1 else len(rows)
H = rows[hi]; ai, ii, si = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
agg, sam, tot, ts, base, prev = collections.Counter(), collections.Counter(), 0, 0, None, "?"
for r in rows[hi + 1:end]:
    if len(r) <= max(ii, si): continue
    try: a = int(r[ai], 16)
    except ValueError: continue
    if base is None: base = a
    ln = line_of.get(a - base)
    if ln and ln[0] == base_name: prev = fn(ln[1])
    agg[prev] += int(r[ii]); sam[prev] += int(r[si]); tot += int(r[ii]); ts += int(r[si])
print("kernel %s: %d warp-instructions, %d stall samples" % (kname, tot, ts))
for f, n in agg.most_common(top):
    print("%5.1f%% inst  %5.1f%% samples  %s" % (100.0 * n / tot, 100.0 * sam[f] / max(1, ts), f))
