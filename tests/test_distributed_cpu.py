"""N > 1 path on CPU: two gloo ranks shard a HyperStudy's combinations (or an OnlineStudy's hypotheses), run them
through the CPU oracle engine and merge (all-gather of evidences, max + sum all-reduce of the running average; for the
online study an all-gather of the evidence increments per step and a sum all-reduce of the mixture).  Result must
equal the golden of the unsharded reference run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out_dir, deal=None):
    import torch.distributed as td
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    import bayesloop_b200 as bl
    import parity
    from bayesloop_b200 import engine
    from conftest import ORACLE_SO
    engine.set_default_engine(engine.Engine(ORACLE_SO, 'cpu'))
    bl.HyperStudy.shareDeal = deal
    S, got = parity.run_case(name, bl)
    shard = S._dev['rows'] if type(S).__name__ == 'OnlineStudy' else S.sweepStats['rows']
    stats = getattr(S, 'sweepStats', {})
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), shard=np.array(shard), deal=str(stats.get('deal')),
             shared=bool(stats.get('shared')), **got)
    td.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(name, tmp_path, world=2, deal=None, keep_stats=False):
    from conftest import build_oracle
    build_oracle()
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path), deal), nprocs=world, join=True)
    ranks = [dict(np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r))) for r in range(world)]
    for r in ranks:
        stats = {k: r.pop(k) for k in ('deal', 'shared')}
        if keep_stats:
            r['stats'] = stats
    return ranks


def test_two_ranks_reproduce_the_golden_hyperstudy(tmp_path):
    import parity
    from conftest import load_golden
    name = 'syn_hyper_poisson_sweep'
    ranks = _run(name, tmp_path)
    want = load_golden(name)
    # 12 widths of growing cost, dealt by predicted cost in laps of alternating direction (distributed.deal_by_cost)
    assert [list(r['shard']) for r in ranks] == [[0, 3, 4, 7, 8, 11], [1, 2, 5, 6, 9, 10]]
    for r in ranks:
        r.pop('shard')
        parity.compare(name, r, want, rtol=2e-9, atol_post=1e-13)


def test_two_ranks_changepoint_study_with_uneven_shards(tmp_path):
    import parity
    from conftest import load_golden
    name = 'ref_cps_coal_all'  # 109 combos -> 55 + 54, dealt round-robin
    ranks = _run(name, tmp_path)
    want = load_golden(name)
    assert [len(r['shard']) for r in ranks] == [55, 54] and ranks[1]['shard'][0] == 1
    for r in ranks:
        r.pop('shard')
        parity.compare(name, r, want, rtol=2e-9, atol_post=1e-13)


def test_two_ranks_online_study_shards_hypotheses(tmp_path):
    """C5 in miniature over two ranks: 6 random-walk pairs + 3 regime-switch values + Independent = 10 hypotheses,
    dealt round-robin; every rank ends up with the complete, identical results."""
    import parity
    from conftest import load_golden
    name = 'syn_online_mixed'
    ranks = _run(name, tmp_path)
    want = load_golden(name)
    assert [list(r['shard']) for r in ranks] == [[0, 2, 4, 6, 8], [1, 3, 5, 7, 9]]
    for r in ranks:
        r.pop('shard')
        parity.compare(name, r, want, rtol=2e-9, atol_post=1e-13)


def test_three_ranks_online_study_with_uneven_shards(tmp_path):
    """The reference's own two-model online test (tests/test_onlinestudy.py:34-84): 4 + 1 hypotheses over three
    ranks (2 + 2 + 1)."""
    import parity
    from conftest import load_golden
    name = 'ref_online_2tm'
    ranks = _run(name, tmp_path, world=3)
    want = load_golden(name)
    assert [list(r['shard']) for r in ranks] == [[0, 3], [1, 4], [2]]
    for r in ranks:
        r.pop('shard')
        parity.compare(name, r, want, rtol=2e-9, atol_post=1e-13)


def test_more_ranks_than_hypotheses(tmp_path):
    """One hypothesis (Static) on two ranks: the second rank owns nothing and still reports the complete results."""
    import parity
    from conftest import load_golden
    name = 'ref_online_static'
    ranks = _run(name, tmp_path)
    want = load_golden(name)
    assert [list(r['shard']) for r in ranks] == [[0], []]
    for r in ranks:
        r.pop('shard')
        parity.compare(name, r, want, rtol=2e-9, atol_post=1e-13)


@pytest.mark.parametrize('world,deal,want_deal,first_rows', [
    (2, None, 'changepoints', [0, 1, 2, 3, 4, 5, 12, 13]),   # the cost model: 6 groups never fill a launch, so deal change-points
    (2, 'groups', 'groups', [0, 2, 4, 6, 8, 10, 12, 14]),
    (3, 'changepoints', 'changepoints', [0, 1, 2, 3, 4, 5, 18, 19]),
    (3, 'groups', 'groups', [0, 3, 6, 9, 12, 15, 18, 21]),
    (5, 'changepoints', 'None', None),  # 8 change-points over 5 ranks: < 2 each -> plain deal (by predicted cost)
])
def test_changepoint_study_with_prefix_sharing_dealt_by_groups_or_by_changepoints(tmp_path, world, deal, want_deal, first_rows):
    """C4 in miniature (8 change-points x 6 groups of random-walk widths) with the change-point prefix sharing under
    torch.distributed: whole groups or whole change-points per rank (core._deal_shared), every rank's merged result
    equal to the golden of the unsharded reference run, every combination owned by exactly one rank."""
    import parity
    from conftest import load_golden
    name = 'syn_cps_gauss_2d'
    ranks = _run(name, tmp_path, world=world, deal=deal, keep_stats=True)
    want = load_golden(name)
    owned = np.sort(np.concatenate([r['shard'] for r in ranks]))
    assert list(owned) == list(range(48))
    if first_rows is not None:
        assert list(ranks[0]['shard'][:8]) == first_rows
    else:
        assert sorted(len(r['shard']) for r in ranks) == [9, 9, 10, 10, 10]
    for r in ranks:
        stats = r.pop('stats')
        assert str(stats['deal']) == want_deal and bool(stats['shared']) == (want_deal != 'None')
        r.pop('shard')
        parity.compare(name, r, want, rtol=2e-9, atol_post=1e-13)
