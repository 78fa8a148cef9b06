"""SASS opcode histogram of the built library, per kernel family (evidence that the hot kernels are Blackwell-native:
UTCIMMA = tcgen05.mma.kind::i8, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk through the TMA engine, UTCBAR = tcgen05.commit,
DMMA = FP64 mma.sync, SYNCS = mbarrier).      python tools/sass_hist.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'beta-cores_b200', 'lib', 'libbetacores.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
fam = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        key = re.sub(r'<.*', '', name.replace('void ', '')).strip()
        cur = fam.setdefault(key, {'n': 0, 'ops': collections.Counter()})
        cur['n'] += 1
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur is not None:
        cur['ops'][m.group(1)] += 1
KEY = ['UTCIMMA', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'DMMA', 'DFMA', 'DADD', 'DMUL', 'IMAD', 'I2F', 'LDL', 'STL', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR', 'USETMAXREG']
print('SASS opcode histogram of %s (cuobjdump -sass; instantiations of a template are summed)' % os.path.relpath(so, ROOT))
print('%-44s %5s %8s  %s' % ('kernel family', 'inst.', 'SASS', ' '.join('%s' % k for k in KEY)))
tot = collections.Counter()
for k, v in fam.items():
    tot.update(v['ops'])
    print('%-44s %5d %8d  %s' % (k[:44], v['n'], sum(v['ops'].values()), ' '.join('%*d' % (len(kk), v['ops'].get(kk, 0)) for kk in KEY)))
print('%-44s %5s %8d  %s' % ('TOTAL', '', sum(tot.values()), ' '.join('%*d' % (len(kk), tot.get(kk, 0)) for kk in KEY)))
print()
for k in ('bc::k_project_q', 'bc::k_project'):
    if k in fam:
        print('%s: top opcodes  %s' % (k, ', '.join('%s %d' % kv for kv in fam[k]['ops'].most_common(24))))
