"""The C-ABI libraries load and export every symbol include/blgrid.h declares (no compute calls, no GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ORACLE_SO, ROOT, build_oracle

HEADER = os.path.join(ROOT, 'include', 'blgrid.h')
CUDA_SO = os.path.join(ROOT, 'bayesloop_b200', 'csrc', 'libblgrid.so')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(blg_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ('blg_version', 'blg_last_error', 'blg_backend', 'blg_plan_create', 'blg_plan_destroy', 'blg_forward',
                 'blg_backward', 'blg_accumulate', 'blg_scale', 'blg_finalize', 'blg_mix', 'blg_launch_count',
                 'blg_plan_set_option'):
        assert must in names


def _ensure_cuda_lib():
    if not os.path.exists(CUDA_SO):
        import sys
        subprocess.check_call([sys.executable, os.path.join(ROOT, '__graft_entry__.py')])
    return CUDA_SO


@pytest.mark.parametrize('which', ['cuda', 'oracle'])
def test_library_exports_every_declared_symbol(which):
    path = _ensure_cuda_lib() if which == 'cuda' else build_oracle()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), '{} does not export {}'.format(os.path.basename(path), name)
    lib.blg_version.restype = ctypes.c_int
    lib.blg_backend.restype = ctypes.c_char_p
    assert lib.blg_version() == 1
    backend = lib.blg_backend().decode()
    assert backend == ('cuda:sm_100a' if which == 'cuda' else 'cpu-oracle')


def test_cuda_library_contains_sm100a_sass_with_bulk_copies():
    """SASS evidence that the hot path is Blackwell-native: sm_100a cubin with UBLKCP (cp.async.bulk) in the
    resident kernels, and no PTX-only fallback."""
    cuobjdump = '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-sass', _ensure_cuda_lib()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert 'UBLKCP' in out, 'bulk-async copy instructions missing from the SASS'
    assert 'fwd_resident_kernel' in out and 'bwd_resident_kernel' in out
    # thread-block-cluster kernels: cluster barrier, st.async into distributed shared memory, mbarrier transactions
    assert 'fwd_cluster2d_kernel' in out and 'bwd_cluster2d_kernel' in out
    for mnemonic in ('UCGABAR_ARV', 'UCGABAR_WAIT', 'STAS.128', 'SYNCS.ARRIVE.TRANS64'):
        assert mnemonic in out, mnemonic + ' missing from the SASS'
    # tiled online step: the cp.async (LDGSTS) and the LDG load variant, and no local memory in either
    assert 'online2d_tile_kernel' in out and 'online2d_finish_kernel' in out
    blocks = out.split('Function : ')
    tiles = [b for b in blocks if b.startswith('_ZN3blg20online2d_tile_kernel')]
    assert len(tiles) == 2
    assert any('LDGSTS' in b for b in tiles)
    for b in tiles:
        assert 'DFMA' in b and ' LDL' not in b and ' STL' not in b, 'register spill in the tile kernel'
    # headline kernels of the bench (C2: grid 1000 = 4 tiles of 64 cells per compute warp, 160 threads): the convolution
    # is issued as FP64 matrix instructions, the backward pass stages its rows with bulk copies, nothing spills
    # (profiles/r2_sass_resources.txt)
    headline = [b for b in blocks if b.startswith(('_ZN3blg21fwd_fast1d_mma_kernelILi4ELi160ELb0E',
                                                   '_ZN3blg21bwd_fast1d_mma_kernelILi4ELi160E'))]
    assert len(headline) == 2
    for b in headline:
        assert 'DMMA' in b, 'FP64 matrix instructions missing from a headline kernel'
        assert 'UBLKCP' in b or 'fwd_fast1d' in b[:40], 'bulk copies missing from the backward headline kernel'
        assert ' LDL' not in b and ' STL' not in b, 'register spill in a headline kernel'


def test_product_fails_loudly_without_cuda(monkeypatch):
    import torch
    from bayesloop_b200 import engine
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    engine.set_default_engine(None)
    with pytest.raises(engine.EngineError):
        engine.default_engine()


def test_ctypes_structs_match_the_header(tmp_path):
    """The Python binding lays its structures out exactly like include/blgrid.h (sizes from a C compiler)."""
    from bayesloop_b200 import engine
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include "blgrid.h"\nint main(void) { printf("%zu %zu %zu %zu\\n", '
                   'sizeof(blg_problem), sizeof(blg_program), sizeof(blg_inputs), sizeof(blg_outputs)); return 0; }\n')
    exe = tmp_path / 'sizes'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    want = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    got = [ctypes.sizeof(engine._Problem), ctypes.sizeof(engine._Program), ctypes.sizeof(engine._Inputs),
           ctypes.sizeof(engine._Outputs)]
    assert got == want


def test_integration_stub_declares_the_same_structures_as_the_binding():
    """The ctypes stub INTEGRATION.md shows a bayesloop maintainer is executable documentation: its four structures must
    have the field names, order and sizes of the binding the product uses (bayesloop_b200/engine.py), which the test
    above ties to include/blgrid.h."""
    import re
    from bayesloop_b200 import engine
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = text[text.index('class Problem(ctypes.Structure):'):text.index('def fit_on_gpu')]
    scope = {'ctypes': ctypes}
    exec(block, scope)  # class definitions only
    for name, ours in (('Problem', engine._Problem), ('Program', engine._Program), ('Inputs', engine._Inputs),
                       ('Outputs', engine._Outputs)):
        stub = scope[name]
        assert [f[0] for f in stub._fields_] == [f[0] for f in ours._fields_], name
        assert ctypes.sizeof(stub) == ctypes.sizeof(ours), name
        assert [getattr(stub, f[0]).offset for f in stub._fields_] == [getattr(ours, f[0]).offset for f in ours._fields_], name
    # every entry point the stub calls is declared by the header
    called = set(re.findall(r'_lib\.(blg_[a-z_]+)', text))
    assert called and called <= set(declared_symbols()), called - set(declared_symbols())


def test_python_constants_match_the_header_enums():
    """Flags, operator codes and observation-model codes are duplicated in the binding (engine.py,
    transitionModels.py, observationModels.py): they must be the values of include/blgrid.h."""
    from bayesloop_b200 import engine, observationModels as om, transitionModels as tm
    text = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    enum = {}
    for name, value in re.findall(r'\b(BLG_[A-Z0-9_]+)\s*=\s*([^,}\n]+)', text):
        value = value.strip()
        m = re.fullmatch(r'1u\s*<<\s*(\d+)', value)
        enum[name] = 1 << int(m.group(1)) if m else int(value)
    for py, c in (('F_EVIDENCE_ONLY', 'BLG_F_EVIDENCE_ONLY'), ('F_INIT_STATE', 'BLG_F_INIT_STATE'),
                  ('F_TRANSITION_FIRST', 'BLG_F_TRANSITION_FIRST'), ('F_SAVE_STATE', 'BLG_F_SAVE_STATE'),
                  ('F_ACCUMULATE', 'BLG_F_ACCUMULATE'), ('F_NORMALIZE_ROWS', 'BLG_F_NORMALIZE_ROWS'),
                  ('F_RAW_ALPHA', 'BLG_F_RAW_ALPHA'), ('F_RAW_POSTERIOR', 'BLG_F_RAW_POSTERIOR'),
                  ('F_SEPARABLE_ROWS', 'BLG_F_SEPARABLE_ROWS')):
        assert getattr(engine, py) == enum[c], (py, c)
    for py, c in (('OP_GRW', 'BLG_OP_GRW'), ('OP_REGIME', 'BLG_OP_REGIME'), ('OP_RESET', 'BLG_OP_RESET'),
                  ('OP_NOTEQUAL', 'BLG_OP_NOTEQUAL')):
        assert getattr(tm, py) == enum[c], (py, c)
    for py, c in (('KIND_POISSON', 'BLG_OM_POISSON'), ('KIND_GAUSSIAN', 'BLG_OM_GAUSSIAN'),
                  ('KIND_SCALED_AR1', 'BLG_OM_SCALED_AR1'), ('KIND_AR1', 'BLG_OM_AR1'),
                  ('KIND_WHITE_NOISE', 'BLG_OM_WHITE_NOISE'), ('KIND_GAUSSIAN_MEAN', 'BLG_OM_GAUSSIAN_MEAN'),
                  ('KIND_LAPLACE', 'BLG_OM_LAPLACE'), ('KIND_BERNOULLI', 'BLG_OM_BERNOULLI'), ('KIND_TABLE', 'BLG_OM_TABLE')):
        assert getattr(om, py) == enum[c], (py, c)
    assert engine.MAX_OPS == int(re.search(r'#define BLG_MAX_OPS (\d+)', text).group(1))
