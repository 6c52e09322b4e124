"""Static SASS instruction counts per kernel of promp_b200/libpromp_b200.so (cuobjdump -sass): which hardware paths each
kernel uses.  usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = collections.OrderedDict([
    ('UTCHMMA', r'\bUTCHMMA'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UTCBAR', r'\bUTCBAR'), ('UBLKCP', r'\bUBLKCP'),
    ('SYNCS', r'\bSYNCS'), ('HMMA.TF32', r'\bHMMA\.[0-9]+\.F32\.TF32'), ('FFMA', r'\bFFMA'), ('DFMA', r'\bDFMA'),
])


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'promp_b200', 'libpromp_b200.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
    print('# SASS evidence (cuobjdump -sass promp_b200/libpromp_b200.so): tcgen05 shows up as UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st),')
    print('# UTCBAR (tcgen05.commit); TMA bulk copies (cp.async.bulk) as UBLKCP with SYNCS (mbarrier) around them; the warp-level tensor-core path')
    print('# (mma.sync m16n8k8 tf32, weight gradients) as HMMA.*.TF32; float64 scans / Gram / solve as DFMA')
    print('# kernel | static instruction counts')
    tot = collections.Counter()
    for fn in sorted(counts):
        c = counts[fn]
        tot.update(c)
        print('%s | %s' % (fn, ' '.join('%s %d' % (k, c[k]) for k in PAT)))
    print('# TOTAL | %s' % ' '.join('%s %d' % (k, tot[k]) for k in PAT))


if __name__ == '__main__':
    main()
