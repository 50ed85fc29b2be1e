#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per kernel the registers, static shared memory, stack,
and the SASS instruction mix that matters for the claims in DESIGN.md (local-memory spills LDL/STL, FP64 matrix
instructions DMMA, FP64 FMA, bulk-async copies UBLKCP, DSMEM stores STAS, mbarrier SYNCS, cluster barriers UCGABAR).

    python tools/sass_resources.py [path/to/libblgrid.so] > profiles/<tag>_sass_resources.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'bayesloop_b200', 'csrc', 'libblgrid.so')
COUNTED = ('LDL', 'STL', 'DMMA', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'UBLKCP', 'STAS', 'SYNCS', 'UCGABAR', 'BAR', 'LDG', 'STG',
           'LDS', 'STS', 'LDGSTS')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r'^void ', '', name)
    name = name.replace('blg::', '')
    return re.sub(r'\(.*$', '', name)


def resources():
    txt = subprocess.run(['cuobjdump', '--dump-resource-usage', LIB], capture_output=True, text=True).stdout
    res, fn = {}, None
    for line in txt.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', line)
        if m and fn:
            res[fn] = tuple(int(x) for x in m.groups())
            fn = None
    return res


def sass_mix():
    proc = subprocess.Popen(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True)
    mix, fn = collections.defaultdict(collections.Counter), None
    for line in proc.stdout:
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            continue
        if fn is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            mix[fn]['total'] += 1
            if op in COUNTED:
                mix[fn][op] += 1
    proc.wait()
    return mix


def main():
    res, mix = resources(), sass_mix()
    names = demangle(sorted(res))
    cols = ('LDL', 'STL', 'DMMA', 'DFMA', 'UBLKCP', 'LDGSTS', 'STAS', 'SYNCS', 'UCGABAR', 'BAR')
    print('# %s (%d bytes, %d kernels) -- cuobjdump --dump-resource-usage + cuobjdump -sass'
          % (os.path.relpath(LIB, ROOT), os.path.getsize(LIB), len(res)))
    print('# dynamic shared memory is set at launch (not shown); SHARED is the static part')
    print('%-78s %4s %5s %6s %7s ' % ('kernel', 'REG', 'STACK', 'SHARED', 'SASS') + ' '.join('%6s' % c for c in cols))
    families = collections.defaultdict(lambda: [0, 0, 0])
    for fn in sorted(res, key=lambda f: short(names[f])):
        reg, stack, shared, local = res[fn]
        c = mix.get(fn, {})
        label = short(names[fn])
        print('%-78s %4d %5d %6d %7d ' % (label[:78], reg, stack, shared, c.get('total', 0))
              + ' '.join('%6d' % c.get(k, 0) for k in cols))
        fam = re.sub(r'<.*$', '', label)
        families[fam][0] += 1
        families[fam][1] += c.get('LDL', 0) + c.get('STL', 0)
        families[fam][2] = max(families[fam][2], reg)
    print()
    print('# per kernel family: instantiations, spill instructions (LDL + STL) summed, largest register count')
    for fam, (n, spills, reg) in sorted(families.items()):
        print('%-40s %3d instantiation(s)  spills %5d  max REG %3d' % (fam, n, spills, reg))


if __name__ == '__main__':
    main()
