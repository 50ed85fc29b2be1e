"""save / load of study objects (reference: bayesloop/fileIO.py:10-37, which pickles the study with dill).

Studies pickle with the standard library: fitted `Study` / `HyperStudy` / `ChangepointStudy` objects hold host arrays
only; an `OnlineStudy` brings its device-resident hypothesis posteriors to the host on the way out and re-creates the
device side lazily after loading (core.py: OnlineStudy.__getstate__ / _resume), so a stream can be checkpointed on one
GPU and resumed on another."""
import pickle


def _pickler():
    """The reference pickles with dill so that studies holding lambdas (priors, NumPy likelihoods) can be saved
    (fileIO.py:4, :21); use it when present, else cloudpickle (same capability, readable by plain pickle.load), else the
    standard library (module-level functions only)."""
    for name in ('dill', 'cloudpickle'):
        try:
            return __import__(name)
        except ImportError:
            continue
    return pickle


def save(filename, study):
    """Write `study` to `filename` (counterpart of bayesloop.save, fileIO.py:10-23)."""
    with open(filename, 'wb') as f:
        _pickler().dump(study, f, protocol=pickle.HIGHEST_PROTOCOL)
    print('+ Successfully saved current study.')


def load(filename):
    """Read a study written by `save` (counterpart of bayesloop.load, fileIO.py:26-37)."""
    with open(filename, 'rb') as f:
        study = _pickler().load(f) if _pickler().__name__ == 'dill' else pickle.load(f)
    print('+ Successfully loaded study.')
    return study
