"""Per-SM timeline of the warp-specialised 1-D kernels from a BLG_TRACE dump (tools/trace_c2.py):
    python tools/sm_timeline.py gpurun_out/r2s_trace
For every kernel: when the SMs finish, and how much SM-time is spent with 4 / 3 / 2 / 1 chains still running on the SM
(a chain that runs alone cannot keep the FP64 pipe busy during its epilogue and barrier)."""
import csv
import glob
import sys

import numpy as np

prefix = sys.argv[1]
R = np.load(prefix + '.radius.npy')
for path in sorted(glob.glob(prefix + '.*.csv')):
    rows = [r for r in csv.DictReader(open(path)) if int(r['end_ns']) > 0 and int(r['combo']) >= 0]
    t0 = min(int(r['start_ns']) for r in rows)
    sms = {}
    for r in rows:
        sms.setdefault(int(r['smid']), []).append((int(R[int(r['combo'])]), (int(r['end_ns']) - t0) / 1e6))
    ends = np.array([max(e for _, e in v) for v in sms.values()])
    first = np.array([min(e for _, e in v) for v in sms.values()])
    active = np.zeros(6)
    for v in sms.values():
        e = sorted(x[1] for x in v)
        prev = 0.0
        for i, x in enumerate(e):
            active[len(e) - i] += x - prev
            prev = x
    print(path.split('/')[-1])
    print('  SMs %d, chains %d; last chain of an SM ends at min / mean / max %.2f / %.2f / %.2f ms; FIRST chain of an SM '
          'ends at mean %.2f ms' % (len(sms), len(rows), ends.min(), ends.mean(), ends.max(), first.mean()))
    tot = active.sum()
    print('  share of SM-time with k chains running: ' + ', '.join('k=%d %.1f %%' % (k, 100 * active[k] / tot)
                                                                   for k in (4, 3, 2, 1)) +
          '; idle until the kernel ends %.1f %%' % (100 * (ends.max() * len(sms) - tot) / (ends.max() * len(sms))))
    for k in sorted(sms)[:4]:
        print('  SM %3d: (radius, end ms) %s' % (k, sorted(sms[k])))
