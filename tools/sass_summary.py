"""Binary evidence for the shipped library: sha256 of libag2v_sm100a.so and, per kernel, the counts of the
SASS mnemonics that prove the Blackwell-native paths (cuobjdump -sass; B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA / bulk copies,
UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, SYNCS = mbarrier, UCGABAR / CGA* = cluster barriers.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ag2video_b200', 'libag2v_sm100a.so')
PATTERNS = ['UTC[A-Z]*MMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'SYNCS', 'UCGABAR', 'MEMBAR', 'ATOM', 'RED']


def main():
    sha = hashlib.sha256(open(LIB, 'rb').read()).hexdigest()
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    try:
        names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True,
                               text=True).stdout.splitlines()
    except Exception:
        names = re.findall(r'Function : (\S+)', sass)
    blocks = re.split(r'\n\s*Function : \S+', sass)[1:]
    print('libag2v_sm100a.so  sha256 %s  (%d bytes, %d kernels; arch %s)' % (
        sha, os.path.getsize(LIB), len(blocks), ', '.join(sorted(set(re.findall(r'arch = (sm_\w+)', sass))))))
    print('%-86s %6s ' % ('kernel', 'instr') + ' '.join('%8s' % p.replace('[A-Z]*', '*') for p in PATTERNS))
    total = [0] * len(PATTERNS)
    for name, body in sorted(zip(names, blocks)):
        lines = [l for l in body.splitlines() if re.search(r'/\*[0-9a-f]{4,}\*/', l)]
        ops = [re.sub(r'^\s*/\*[0-9a-f]+\*/\s*(@!?U?P\d+\s+)?', '', l).split('(')[0].split()[0] if l.strip() else '' for l in lines]
        counts = [sum(1 for o in ops if re.match(p + r'(\.|$|\s|;)', o)) for p in PATTERNS]
        total = [a + b for a, b in zip(total, counts)]
        short = re.sub(r'\(.*', '', name)[:86]
        print('%-86s %6d ' % (short, len(ops)) + ' '.join('%8d' % c for c in counts))
    print('%-86s %6s ' % ('TOTAL', '') + ' '.join('%8d' % c for c in total))


if __name__ == '__main__':
    sys.exit(main())
