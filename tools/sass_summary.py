"""SASS mnemonic counts per kernel of the built library (evidence of tcgen05 / TMEM / TMA use):
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-aspect-category-sentiment-analysis_b200", "libfcmf_b200.so")
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'HMMA', 'REDG', 'MUFU.EX2', 'MUFU.TANH', 'FFMA2', 'FADD2', 'FMUL2']
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None or '/*' not in line:
        continue
    for k in KEYS:
        if re.search(r'(?<![A-Z0-9_.])' + re.escape(k) + r'(?![A-Z0-9_])', line):
            cnt[cur][k] += 1
print("# SASS mnemonic counts per kernel of libfcmf_b200.so (cuobjdump -sass, sm_100a). UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG/UTMASTG = TMA tensor load/store (cp.async.bulk.tensor), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async,")
print("# HMMA = legacy mma.sync (none expected), FFMA2/FADD2/FMUL2 = packed fp32x2. Kernels using none of the tensor/TMA paths are omitted.")
print("%-100s %s" % ("kernel", " ".join("%9s" % k for k in KEYS)))
for f, c in cnt.items():
    if not any(c[k] for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'HMMA')):
        continue
    name = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(.*', '', name).replace('fcmf::', '')[:98]
    print("%-100s %s" % (name, " ".join("%9d" % c[k] for k in KEYS)))
