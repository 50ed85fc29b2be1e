"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on every case of tests/cases.py.

Run in the build container (the GPU box has no /root/reference):

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py ref_tm_grw # selected cases

The fixtures are committed; the tests only ever read them.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_shim  # noqa: E402
import cases  # noqa: E402


def main(argv):
    bl = ref_shim.import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    names = argv or list(cases.CASES)
    for name in names:
        t0 = time.time()
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all="ignore"):
            S = cases.CASES[name](bl)
            res = cases.extract(S)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **res)
        pinned = cases.REFERENCE_PINNED_LOGE.get(name)
        msg = ""
        if pinned is not None:
            msg = " pinned-diff=%.2e" % abs(float(res["logEvidence"]) - pinned)
        if name in cases.REFERENCE_PINNED_LOG10E:
            msg = " pinned-log10-diff=%.2e" % abs(float(res["logEvidence"]) / np.log(10) -
                                                  cases.REFERENCE_PINNED_LOG10E[name])
        print("%-36s logE=%+.15g  %.2fs%s" % (name, float(res["logEvidence"]), time.time() - t0, msg))


if __name__ == "__main__":
    main(sys.argv[1:])
