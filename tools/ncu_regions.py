"""Stall-sample shares of a kernel's regions between CTA barriers, from the source page of an `ncu --set full
--import-source on` report:  python tools/ncu_regions.py gpurun_out/<report>.ncu-rep [kernel-substring] [top]

Prints, per kernel, the share of warp-stall samples that fall between consecutive BAR.SYNC instructions (the phases of
the barrier-separated kernels of this repo) and the instructions with the most samples."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ''
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for b in blocks:
        if want not in b['name'] or len(b['rows']) < 2:
            continue
        hdr = b['rows'][0]
        isrc, iss = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
        data = [(r[isrc].strip(), int(r[iss] or 0)) for r in b['rows'][1:] if len(r) > iss]
        tot = sum(d[1] for d in data) or 1
        print('==', b['name'], '-- %d instructions, %d stall samples' % (len(data), tot))
        prev = 0
        bars = [i for i, d in enumerate(data) if 'BAR.SYNC' in d[0] or 'BAR.ARV' in d[0]]
        for e in bars + [len(data) - 1]:
            seg = data[prev:e + 1]
            share = 100.0 * sum(x[1] for x in seg) / tot
            if share >= 0.5:
                kinds = {}
                for x in seg:
                    op = x[0].split()[1] if x[0].startswith('@') and len(x[0].split()) > 1 else x[0].split()[0]
                    kinds[op.split('.')[0]] = kinds.get(op.split('.')[0], 0) + 1
                mix = ', '.join('%s %d' % kv for kv in sorted(kinds.items(), key=lambda kv: -kv[1])[:4])
                print('  instr %5d-%5d: %5.1f %% of samples   (%s)' % (prev, e, share, mix))
            prev = e + 1
        print('  top instructions:')
        for i, d in sorted(enumerate(data), key=lambda x: -x[1][1])[:top]:
            print('    %5d  %5.1f %%  %s' % (i, 100.0 * d[1] / tot, d[0][:80]))


if __name__ == '__main__':
    main()
