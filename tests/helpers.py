"""Helpers shared by the oracle / GPU tests: turn a configured product study into (a) raw ABI inputs for an
Engine and (b) an oracle/np_oracle.Problem, from ONE lowering."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def lowered(study):
    """(ops, hyperPrior) of a HyperStudy/Study whose data and models are set (no fit needed)."""
    study._formatData()
    if hasattr(study, '_createHyperGrid'):
        study._createHyperGrid(silent=True)
        hyper = np.asarray(study.hyperGridValues, dtype=float)
        hp = np.asarray(study.flatHyperPriorValues, dtype=float)
    else:
        values = study._unpackAllHyperParameters()
        hyper = np.array([values], dtype=float).reshape(1, len(values))
        hp = np.ones(1)
    ctx = study._lower(hyper, study.formattedTimestamps)
    return ctx.ops, hp, ctx.usesReset


def np_problem(study):
    import np_oracle
    om = study.observationModel
    prior = np.array(study._computePrior(silent=True), dtype=float)
    p = om.prior
    if callable(p):
        base = np.asarray(p(*study.grid), dtype=float) * np.ones(study.gridSize)
    elif isinstance(p, np.ndarray):
        base = np.array(p, dtype=float)
    else:
        base = np.ones(study.gridSize)
    base = base / base.sum()
    return np_oracle.Problem(study.marginalGrid, study.latticeConstant, om.deviceKind, om.segmentLength,
                             study.rawData, prior, reset_base=base)


def online_ops(study):
    """Concatenated per-hypothesis program of an OnlineStudy (what OnlineStudy._setupDevice hands to the engine), as
    host arrays: ops with window/param/radius rows for ALL hypotheses."""
    ops, row, H = [], 0, study.tmCount
    for tm, rows, count in zip(study.transitionModels, study.hyperParameterValues, study.tmCounts):
        hyper = np.asarray(rows, dtype=float).reshape(count, -1) if len(rows) > 0 else np.zeros((1, 0))
        study.setTransitionModel(tm, silent=True)
        ctx = study._lower(hyper, np.array([-1.]), online=True, model=tm)
        for op in ctx.ops:
            full = dict(kind=op['kind'], axis=op['axis'], param=np.zeros(H), radius=np.zeros(H, dtype=np.int32),
                        window=np.zeros((H, 4), dtype=np.int32))
            full['param'][row:row + count] = op['param']
            full['radius'][row:row + count] = op['radius']
            full['window'][row:row + count] = op['window']
            ops.append(full)
        row += count
    return ops


def np_problem_online(study, series):
    """np_oracle.Problem of an OnlineStudy fed with `series` (grid and prior from the study's observation model)."""
    import np_oracle
    om = study.observationModel
    prior = np.array(study._computePrior(silent=True), dtype=float)
    p = om.prior
    if callable(p):
        base = np.asarray(p(*study.grid), dtype=float) * np.ones(study.gridSize)
    elif isinstance(p, np.ndarray):
        base = np.array(p, dtype=float)
    else:
        base = np.ones(study.gridSize)
    base = base / base.sum()
    return np_oracle.Problem(study.marginalGrid, study.latticeConstant, om.deviceKind, om.segmentLength,
                             np.asarray(series, dtype=float), prior, reset_base=base)


def abi_sweep(engine, study, forwardOnly=False, evidenceOnly=False):
    """Run the product's sweep on `engine` and bring everything back to the host."""
    study._engineOverride = engine
    study._formatData()
    study._createHyperGrid(silent=True)
    sw = study._prepareSweep(forwardOnly, evidenceOnly)
    eng, logE, alive, localEv, avg, means = study._executeSweep(sw)
    out = dict(logE=logE, alive=alive, localEvidence=eng.to_host(localEv), local=eng.to_host(sw['local']))
    if avg is not None:
        out['avg'] = eng.to_host(avg)
        out['means'] = eng.to_host(means)
    return out
