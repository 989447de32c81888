"""Join an ncu SASS source page (csv) with nvdisasm -g line info: instructions executed and stall samples per
CUDA source line.  usage: line_profile.py <ncu_source.csv> <nvdisasm.txt> <kernel-substring> [top]"""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offset -> line
line_of = {}
cur = None; infunc = False
for l in open(dis):
    if l.startswith('.text.') or '.section' in l and '.text.' in l:
        infunc = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and infunc: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
H = rows[hi]
ai, ii, si = H.index('Address'), H.index('Instructions Executed'), H.index('# Samples')
base = None
inst = collections.Counter(); samp = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= ii: continue
    a = int(r[ai], 16)
    if base is None: base = a
    ln = line_of.get(a - base)
    inst[ln] += int(r[ii]); samp[ln] += int(r[si])
tot = sum(inst.values()); ts = sum(samp.values())
print('total warp-instructions', tot, 'samples', ts)
for ln, n in inst.most_common(top):
    print('%5.1f%% inst  %5.1f%% samples  %s' % (100.0 * n / tot, 100.0 * samp[ln] / max(1, ts), ln))
