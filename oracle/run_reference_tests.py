"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Runs the reference's OWN test files for the hot
path -- unmodified, straight from /root/reference/tests -- against the product package: `import bayesloop` inside those
files resolves to `bayesloop_b200`, driven through the C ABI into the CPU oracle (no GPU here; on a B200 box pass
--cuda to use libblgrid.so, the reference tree has to be copied there by hand).

    python oracle/run_reference_tests.py            # -> 46 passed, 3 failed (the three out-of-scope transition models)

The parser and plot suites (tests/test_parser.py, tests/test_plot.py) are outside the engine's scope.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = '/root/reference/tests'
FILES = ['test_study.py', 'test_hyperstudy.py', 'test_changepointstudy.py', 'test_onlinestudy.py',
         'test_observationmodels.py', 'test_transitionmodels.py', 'test_fileio.py']

CONFTEST = '''
import sys
sys.path.insert(0, {root!r})
from bayesloop_b200 import engine
if {cuda!r}:
    engine.set_default_engine(engine.Engine(engine.library_path(), 'cuda:0'))
else:
    engine.set_default_engine(engine.Engine({oracle!r}, 'cpu'))
import bayesloop_b200
sys.modules['bayesloop'] = bayesloop_b200
'''


def main():
    cuda = '--cuda' in sys.argv
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'libblgrid_oracle.so'])
    with tempfile.TemporaryDirectory() as tmp:
        for f in FILES:
            shutil.copy(os.path.join(REF_TESTS, f), tmp)  # copies live only in the temporary directory
        with open(os.path.join(tmp, 'conftest.py'), 'w') as f:
            f.write(CONFTEST.format(root=ROOT, cuda=cuda, oracle=os.path.join(ROOT, 'oracle', 'libblgrid_oracle.so')))
        return subprocess.call([sys.executable, '-m', 'pytest', tmp, '-q', '-p', 'no:cacheprovider'], cwd=tmp)


if __name__ == '__main__':
    sys.exit(main())
