"""bayesloop_b200 -- B200-native grid forward-backward inference engine with the bayesloop API.

    import bayesloop_b200 as bl
    S = bl.HyperStudy()
    S.loadData(counts)
    S.set(bl.om.Poisson('rate', bl.oint(0, 12, 1000)),
          bl.tm.GaussianRandomWalk('sigma', bl.cint(0, 0.2, 512), target='rate'))
    S.fit()

Same names as the reference package (bayesloop/__init__.py:4-18) for everything on the hot path; the probability
parser, plotting and Jeffreys-prior derivation of the reference are outside this engine's scope
(DESIGN.md "Out of scope").
"""
from . import observationModels
from . import observationModels as om
from . import transitionModels
from . import transitionModels as tm
from .core import ChangepointStudy, HyperStudy, OnlineStudy, Study
from .exceptions import ConfigurationError, PostProcessingError
from .fileIO import load, save
from .helper import cint, oint


def _out_of_scope(name, where):
    def stub(*args, **kwargs):
        raise NotImplementedError('bayesloop_b200.{} is not part of this engine ({} of the reference is host-side '
                                  'post-processing / symbolic set-up outside the accelerated path; DESIGN.md "Out of '
                                  'scope").'.format(name, where))
    stub.__name__ = name
    return stub


Parser = _out_of_scope('Parser', 'bayesloop/parser.py')
getJeffreysPrior = _out_of_scope('getJeffreysPrior', 'bayesloop/jeffreys.py:17-68')
computeJeffreysPriorAR1 = _out_of_scope('computeJeffreysPriorAR1', 'bayesloop/jeffreys.py:71-108')

__all__ = ['Study', 'HyperStudy', 'ChangepointStudy', 'OnlineStudy', 'observationModels', 'om', 'transitionModels',
           'tm', 'cint', 'oint', 'ConfigurationError', 'PostProcessingError', 'save', 'load']
__version__ = '0.1.0'
