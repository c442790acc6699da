"""cpnest_b200: B200-native nested sampling with the API of johnveitch/cpnest.

    import cpnest_b200 as cpnest
    import cpnest_b200.model
    work = cpnest.CPNest(model, nlive=1000, maxmcmc=1000, verbose=1)
    work.run()

The sampling hot path (constrained MCMC replacement of live points, worst-point selection,
evidence accumulation) runs in hand-written CUDA kernels for sm_100a behind the C-ABI declared
in include/cpnest_b200.h; there is no CPU fallback.
"""
import logging

from .utils import LEVELS, StreamHandler

# the logger tree keeps the reference's name so existing logging configuration applies
logger = logging.getLogger("cpnest")
logger.setLevel(LEVELS[-1])
logger.addHandler(logging.NullHandler())
console_handler = StreamHandler(verbose=0)
logger.addHandler(console_handler)

__version__ = "0.1.0"


def __getattr__(name):
    # CPNest needs the CUDA library; import it lazily so that `import cpnest_b200.nest2pos` etc. work
    # in environments that only post-process.
    if name == "CPNest":
        from .cpnest import CPNest
        return CPNest
    raise AttributeError(name)


__all__ = ["model", "NestedSampling", "parameter", "cpnest", "nest2pos", "proposal", "plot", "logger", "CPNest"]
