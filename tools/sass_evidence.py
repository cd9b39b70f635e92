#!/usr/bin/env python
"""SASS evidence (B200_PROFILING.md "What proves a Blackwell-native kernel"): per-kernel counts of the tensor-core,
tensor-memory, TMA / bulk-copy, cp.async and mbarrier instructions in the shipped library.

    python tools/sass_evidence.py > profiles/r02_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'sgdml_b200', 'libsgdml_b200.so')
PATTERNS = [
    ('DMMA', r'\bDMMA\.'),          # mma.sync.m8n8k4.f64 (FP64 tensor pipe; tcgen05 has no f64 kind)
    ('UTCIMMA', r'\bUTCIMMA\b'),    # tcgen05.mma kind::i8
    ('UTC*MMA(other)', r'\bUTC(H|Q|O)MMA\b'),
    ('LDTM', r'\bLDTM\b'),          # tcgen05.ld
    ('UTCBAR', r'\bUTCBAR\b'),      # tcgen05.commit
    ('UTMALDG', r'\bUTMALDG\b'),    # cp.async.bulk.tensor
    ('UBLKCP', r'\bUBLKCP\b'),      # cp.async.bulk (1-D)
    ('LDGSTS', r'\bLDGSTS\b'),      # cp.async
    ('SYNCS', r'\bSYNCS\b'),        # mbarrier
    ('HMMA/IMMA', r'\b(HMMA|IMMA)\b'),
]


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    fn = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r'\(.*', '', fn)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[fn][name] += 1
    names = [n for n, _ in PATTERNS]
    print('# %s (sm_100a SASS, cuobjdump -sass); columns: %s' % (os.path.basename(LIB), ', '.join(names)))
    print('%-72s %s' % ('kernel', ' '.join('%8s' % n[:8] for n in names)))
    tot = collections.Counter()
    for fn, c in counts.items():
        if not sum(c.values()):
            continue
        tot.update(c)
        print('%-72s %s' % (fn[:72], ' '.join('%8d' % c[n] for n in names)))
    print('%-72s %s' % ('TOTAL', ' '.join('%8d' % tot[n] for n in names)))
    print('# FP64 contractions: DMMA (mma.sync m8n8k4.f64 is the only FP64 tensor instruction of sm_100a; tcgen05 has no f64 kind).')
    print('# Cholesky trailing update on the 5th-generation tensor cores: UTCIMMA (tcgen05.mma kind::i8, int32 accumulators in')
    print('# tensor memory read back with LDTM) through error-free int8 slicing, csrc/ozaki.cu.')


if __name__ == '__main__':
    main()
