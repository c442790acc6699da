"""Post-processing of nested samples: weights, posterior resampling, autocorrelation length.
Host-side numpy.  Same functions and results as cpnest/nest2pos.py:17-199,
written vectorised; pinned against the reference's outputs in tests/test_host_api.py.
"""
import logging

import numpy as np
from numpy import logaddexp

LOGGER = logging.getLogger("cpnest.nest2pos")


def logsubexp(x, y):
    """log(exp(x) - exp(y)) for x >= y (cpnest/nest2pos.py:17-31)."""
    assert np.all(x >= y), "cannot take log of negative number {0!s} - {1!s}".format(str(x), str(y))
    return x + np.log1p(-np.exp(y - x))


def log_integrate_log_trap(log_func, log_support):
    """log of the trapezoid integral of exp(log_func) over exp(log_support) (nest2pos.py:33-42)."""
    log_func = np.asarray(log_func)
    log_support = np.asarray(log_support)
    mids = logaddexp(log_func[:-1], log_func[1:]) - np.log(2)
    widths = logsubexp(log_support[:-1], log_support[1:])
    return logaddexp.reduce(mids + widths)


def compute_weights(data, Nlive):
    """(log_ev, log_wts) of a nested-sampling logL sequence whose last Nlive entries are the final
    live points (nest2pos.py:45-71): volumes shrink by log1p(-1/Nlive) per dead point, then by
    log1p(-1/(Nlive-i)) through the final live points."""
    data = np.asarray(data, dtype=np.float64)
    ndead = len(data) - Nlive
    step = np.log1p(-1.0 / Nlive)
    # volumes enclosing: the dummy first sample, the ndead dead points ...
    vols_start = step * np.arange(ndead + 2)
    # ... and the Nlive final points, each removed from a live set that is one smaller
    shr = np.log1p(-1.0 / (Nlive - np.arange(Nlive - 1)))
    v0 = vols_start[-1] + step
    vols_end = np.concatenate(([v0], v0 + np.cumsum(shr)))
    log_vols = np.concatenate((vols_start, vols_end))
    log_likes = np.concatenate(([-np.inf], data[:ndead], data[ndead:], [data[-1]]))
    log_ev = log_integrate_log_trap(log_likes, log_vols)
    log_dX = logsubexp(log_vols[:-1], log_vols[1:])
    log_wts = log_likes[1:-1] + log_dX[:-1] - log_ev
    return log_ev, log_wts


def compute_weights_from_volumes(log_likes, log_vols):
    """(log_ev, log_wts) like compute_weights, for a run whose enclosed prior volumes are KNOWN: log_vols[0] = 0 belongs to
    the dummy first sample, log_vols[i] to nested sample i (len(log_vols) == len(log_likes) + 1).

    compute_weights (nest2pos.py:45-71) reconstructs the volumes assuming one replacement per iteration at constant
    Nlive.  A device batch removes its K worst points before any replacement arrives, so the j-th of them shrinks the
    volume by 1/(N-j) (the rule the reference itself applies in its final sweep, NestedSampling.py:474-476): over a run
    the assumed and the true volumes drift apart by a factor ln(N/(N-K)) / (K/N) in the exponent, which would over-weight
    the late, high-likelihood samples.  The integrator's own history (NS.state.log_vols) is the consistent choice."""
    log_likes = np.asarray(log_likes, dtype=np.float64)
    log_vols = np.asarray(log_vols, dtype=np.float64)
    assert len(log_vols) == len(log_likes) + 1 and log_vols[0] == 0.0
    log_ev = log_integrate_log_trap(np.concatenate(([-np.inf], log_likes)), log_vols)
    log_dX = logsubexp(log_vols[:-1], log_vols[1:])
    return log_ev, log_likes + log_dX - log_ev


def draw_posterior(data, log_wts, verbose=False):
    """Rejection-sample rows of `data` with probability exp(log_wt - max) (nest2pos.py:73-81)."""
    log_wts = np.asarray(log_wts)
    keep = (log_wts - log_wts.max()) > np.log(np.random.uniform(size=len(log_wts)))
    return data[keep]


def draw_posterior_many(datas, Nlives, verbose=False, log_vols=None):
    """Combine several runs, each weighted by its evidence (nest2pos.py:83-111).  log_vols (not in the reference): per
    run, the recorded log-volumes (compute_weights_from_volumes) or None for the reference's reconstruction."""
    if log_vols is None:
        log_vols = [None] * len(datas)
    pairs = [compute_weights(d["logL"], n) if lv is None else compute_weights_from_volumes(d["logL"], lv)
             for d, n, lv in zip(datas, Nlives, log_vols)]
    log_evs = [p[0] for p in pairs]
    LOGGER.critical("Computed log_evidences: {0!s}".format(str(tuple(log_evs))))
    top = max(log_evs)
    fracs = [np.exp(le - top) for le in log_evs]
    LOGGER.critical("Relative weights of input files: {0!s}".format(str(fracs)))
    per = [f / len(d) for f, d in zip(fracs, datas)]
    fracs = [p / max(per) for p in per]
    LOGGER.critical("Relative weights of input files taking into account their length: {0!s}".format(str(fracs)))
    posts = [draw_posterior(d, p[1]) for d, p in zip(datas, pairs)]
    LOGGER.critical("Number of input samples: {0!s}".format(str([len(p[1]) for p in pairs])))
    LOGGER.critical("Expected number of samples from each input file {0!s}".format(
        str([int(f * len(p)) for f, p in zip(fracs, posts)])))
    kept = [post[np.random.uniform(size=len(post)) < f] for post, f in zip(posts, fracs)]
    result = np.concatenate(kept, axis=None)
    LOGGER.critical("Samples produced: {0:d}".format(result.shape[0]))
    return result


def draw_N_posterior(data, log_wts, N, verbose=False):
    """N weighted draws with replacement (nest2pos.py:113-128)."""
    if N == 0:
        return []
    log_cum = np.concatenate(([-np.inf], logaddexp.accumulate(np.asarray(log_wts))))
    us = np.log(np.random.uniform(size=N))
    return data[np.digitize(us, log_cum) - 1]


def draw_N_posterior_many(datas, Nlives, Npost, verbose=False):
    """nest2pos.py:130-142"""
    pairs = [compute_weights(d["logL"], n) for d, n in zip(datas, Nlives)]
    total = logaddexp.reduce([p[0] for p in pairs])
    Ns = [int(round(Npost * np.exp(p[0] - total))) for p in pairs]
    posts = [draw_N_posterior(d, p[1], n) for d, p, n in zip(datas, pairs, Ns)]
    return np.vstack(posts).flatten()


def autocorrelation(x):
    """FFT autocorrelation, circular, normalised by the variance (nest2pos.py:175-187)."""
    x = np.asarray(x, dtype=np.float64)
    xp = x - x.mean()
    f = np.fft.fft(xp)
    return np.fft.ifft(f.conjugate() * f).real / np.var(x) / len(x)


def acl(x, tolerance=0.01):
    """1 + 2 * sum of the autocorrelation up to its first drop below `tolerance` (nest2pos.py:189-199)."""
    acf = autocorrelation(x)
    below = np.nonzero(~(acf > tolerance))[0]
    n = below[0] if len(below) else len(acf)
    T = 1
    for v in acf[:n]:
        T += 2 * v
    return T


def resample_mcmc_chain(chain, verbose=False, burnin=False):
    """Thin an MCMC chain by its autocorrelation length and redraw it against the
    Metropolis-Hastings rule (nest2pos.py:144-173)."""
    LOGGER.critical("Number of input samples: {0!s}".format(chain.shape[0]))
    if burnin:
        chain = chain[chain.shape[0] // 2 - 1:]
    acls = [acl(chain[n]) for n in chain.dtype.names if n not in ("logL", "logPrior")]
    ACL = int(np.round(np.max(acls)))
    LOGGER.critical("Measured autocorrelation length {0!s}".format(str(ACL)))
    chain = chain[::ACL]
    np.random.shuffle(chain)
    logpost = chain["logL"] + chain["logPrior"]
    keep = np.concatenate(([True], (logpost[1:] - logpost[:-1]) > np.log(np.random.uniform(size=len(chain) - 1))))
    output = chain[keep]
    LOGGER.critical("Returned number of samples {0!s}".format(str(output.shape[0])))
    return output
