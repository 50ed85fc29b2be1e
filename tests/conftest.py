"""Test configuration.

Markers
    gpu      needs a B200 (run by the driver with `-m gpu` on the GPU box); everything else runs on CPU.

Engines
    `oracle_engine`  Engine around oracle/libblgrid_oracle.so (host pointers).  TEST-ONLY seam: lets the CPU suite
                     drive the product's host logic (lowering, hyper-grid, averaging, sharding) without a GPU and
                     is the checker the GPU tests compare the CUDA library against.
    `cuda_engine`    the product: bayesloop_b200/csrc/libblgrid.so on cuda:0 (gpu-marked tests only).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE_SO = os.path.join(ROOT, 'oracle', 'libblgrid_oracle.so')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (sm_100a); run on the B200 box with -m gpu')


def build_oracle():
    src = os.path.join(ROOT, 'oracle', 'blgrid_oracle.c')
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'libblgrid_oracle.so'],
                              stdout=subprocess.DEVNULL)
    return ORACLE_SO


@pytest.fixture(scope='session')
def oracle_engine():
    from bayesloop_b200.engine import Engine
    return Engine(build_oracle(), 'cpu')


@pytest.fixture(scope='session')
def cuda_engine():
    import torch
    from bayesloop_b200.engine import Engine, library_path
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return Engine(library_path(), 'cuda:0')


@pytest.fixture
def use_oracle(oracle_engine):
    from bayesloop_b200 import engine
    engine.set_default_engine(oracle_engine)
    yield oracle_engine
    engine.set_default_engine(None)


@pytest.fixture
def use_cuda(cuda_engine):
    from bayesloop_b200 import engine
    engine.set_default_engine(cuda_engine)
    yield cuda_engine
    engine.set_default_engine(None)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: z[k] for k in z.files}
