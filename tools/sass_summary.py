#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, plus legacy HMMA (must be 0), registers and spills.

    python tools/sass_summary.py [libsgr.so] > profiles/sass_summary.md      (no GPU needed: cuobjdump -sass / -res-usage)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'stylegan_directions_face_reenactment_b200', 'libsgr.so')
PAT = collections.OrderedDict([('UTC*MMA', r'\bUTC\w*MMA'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UTMALDG', r'\bUTMALDG'),
                               ('UTMASTG', r'\bUTMASTG'), ('UBLKCP', r'\bUBLKCP'), ('UTCBAR', r'\bUTCBAR'), ('SYNCS', r'\bSYNCS'),
                               ('SHFL', r'\bSHFL'), ('HMMA', r'\bHMMA'), ('STL/LDL (spill)', r'\b(STL|LDL)\b')])


def demangle(names):
    out = subprocess.run(['cu++filt'] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    out = [re.sub(r'\((int|bool|unsigned int)\)', '', o) for o in out]
    return [re.sub(r'^void ', '', re.sub(r'\(.*$', '', o)).replace('sgr::', '') for o in out]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for ln in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        for k, p in PAT.items():
            if re.search(p, ln):
                counts[cur][k] += 1
    regs = {}
    fn = None
    for ln in res.splitlines():
        m = re.match(r'\s*Function (\S+):', ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', ln)
        if m and fn:
            loc = re.search(r'LOCAL:(\d+)', ln)
            stk = re.search(r'STACK:(\d+)', ln)
            regs[fn] = (int(m.group(1)), max(int(loc.group(1)) if loc else 0, int(stk.group(1)) if stk else 0))
    names = demangle(order)
    print('# SASS evidence per kernel of libsgr.so (`cuobjdump -sass`, `-res-usage`; tools/sass_summary.py)\n')
    print('sm_100a only.  `UTC*MMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG/UTMASTG` = TMA tensor load/store, `UBLKCP` = '
          'cp.async.bulk, `HMMA` = legacy mma.sync (none).  LOCAL = bytes of local memory / stack per thread (register spills, local arrays).\n')
    print('| kernel | ' + ' | '.join(PAT) + ' | regs | LOCAL |')
    print('|---|' + '---|' * (len(PAT) + 2))
    tot = collections.Counter()
    for mangled, name in sorted(zip(order, names), key=lambda t: t[1]):
        c = counts[mangled]
        tot.update(c)
        r = regs.get(mangled, ('?', '?'))
        print('| `%s` | ' % name[:100] + ' | '.join(str(c[k]) for k in PAT) + ' | %s | %s |' % r)
    print('| **total** | ' + ' | '.join(str(tot[k]) for k in PAT) + ' | | |')


if __name__ == '__main__':
    main()
