#!/usr/bin/env python
"""Counts the SASS mnemonics that prove the tcgen05 / TMEM / TMA path per kernel of librecoder_b200.so
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, SYNCS = mbarrier).  Writes profiles/sass_summary.txt.  Needs only cuobjdump (no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'recoder_b200', 'csrc', 'librecoder_b200.so')
PATTERNS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'MUFU.EX2',
            'LDG.E.128', 'STG.E.128', 'RED', 'ATOM', 'MULTIMEM']


def main():
  out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  per = collections.OrderedDict()
  cur = None
  for ln in out.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
      name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
      cur = per.setdefault(name.split('(')[0], collections.Counter())
      continue
    if cur is None:
      continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m:
      cur['_instructions'] += 1
      op = m.group(1)
      for p in PATTERNS:
        if op.startswith(p):
          cur[p] += 1
  lines = ['# SASS evidence per kernel of recoder_b200/csrc/librecoder_b200.so (cuobjdump -sass, sm_100a)',
           '# columns: instructions | ' + ' '.join(PATTERNS), '']
  for name, c in per.items():
    marks = ' '.join('%s=%d' % (p, c[p]) for p in PATTERNS if c[p])
    lines.append('%-70s %6d  %s' % (name[:70], c['_instructions'], marks))
  tot = collections.Counter()
  for c in per.values():
    tot.update(c)
  lines += ['', 'TOTAL ' + ' '.join('%s=%d' % (p, tot[p]) for p in PATTERNS if tot[p])]
  path = os.path.join(ROOT, 'profiles', 'sass_summary.txt')
  with open(path, 'w') as fh:
    fh.write('\n'.join(lines) + '\n')
  print('\n'.join(l for l in lines if 'UTCHMMA' in l or l.startswith('TOTAL')))


if __name__ == '__main__':
  sys.exit(main())
