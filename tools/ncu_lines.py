"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS rows with sample counts) with the line
table nvdisasm prints for the cubin (needs -lineinfo).   python tools/ncu_lines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [top]"""
import csv, io, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[h]
si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
sass = [(r[1].strip(), int(r[si] or 0), int(r[ii] or 0)) for r in rows[h+1:] if len(r) > si]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# walk the function's listing: remember the current line annotation, collect instructions in order
infn, cur, lines = False, None, []
for l in dis.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append(cur)
if len(lines) != len(sass):
    print('warning: %d SASS rows in the report, %d in the listing' % (len(sass), len(lines)))
agg = collections.Counter(); inst = collections.Counter()
for (txt, smp, ex), ln in zip(sass, lines):
    agg[ln] += smp; inst[ln] += ex
tot = sum(agg.values())
print('%s: %d samples' % (kname, tot))
src = {}
for (f, n), c in agg.most_common(top):
    if f not in src:
        try: src[f] = open('/root/repo/beta-cores_b200/csrc/' + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][n-1].strip() if 0 < n <= len(src[f]) else ''
    print('%5.1f%% %7d inst  %s:%d  %s' % (100.*c/max(tot, 1), inst[(f, n)], f, n, text[:110]))
