"""CPNest: the user-facing facade, drop-in for `cpnest.CPNest` (cpnest/cpnest.py:33-550).

Same constructor arguments, same attributes after `run()` (`NS.logZ`, `NS.state.info`,
`nested_samples`, `posterior_samples`, ...) and the same output files.  What differs is where the
work happens: the reference forks `nthreads` sampler processes and runs the nested-sampling loop
in the parent (cpnest/cpnest.py:181-285); here one device context evolves a whole batch of
constrained chains per kernel launch and keeps the live set, the evidence integrator and the
termination test on the GPU.
"""
import logging
import os
import pickle
import signal
import sys
import time

import numpy as np

from .engine import Engine
from .NestedSampling import LivePointList, NestedSampler
from .parameter import LivePoint
from .utils import LogFile

LOGGER = logging.getLogger("cpnest.cpnest")


class CheckPoint(Exception):
    pass


def sighandler(signal, frame):
    # cpnest/cpnest.py:27-30: the handlers installed with resume=True raise, run() checkpoints and exits with 130
    raise CheckPoint()


class CPNest(object):
    """
    cp = CPNest(usermodel, nlive=100, output='./', verbose=0, seed=None, maxmcmc=100, nthreads=None, ...)

    usermodel : cpnest_b200.model.Model with a `device_functor`
    nlive     : number of live points
    poolsize  : kept for API compatibility; the proposal ensemble is the live set itself
    maxmcmc   : maximum chain length
    nthreads  : in the reference, the number of sampler processes == points replaced per iteration
                (cpnest/NestedSampling.py:324); here a lower bound for the batch size, see `batch`
    nslice    : number of slice samplers.  nslice == nthreads (as in the reference's own tests and examples,
                tests/test_2d_slice.py:30, examples/test_50d.py:37) runs the slice kernel alone.  A mixed population
                (the reference spreads the kinds over its processes, cpnest/cpnest.py:197-261) splits the chains of every
                batch in the proportions (nthreads - nslice - nhamiltonian) : nslice : nhamiltonian; each kind keeps its
                own chain-length controller and the kinds' kernels run side by side
    nhamiltonian : number of HMC samplers; nhamiltonian == nthreads runs the constrained-leapfrog kernel alone
                (tests/test_2d_hmc.py:30).  The reflection normal is the functor's analytic gradient
    proposals : dict(mhs=ProposalCycle subclass, sli=EnsembleSliceProposalCycle subclass, hmc=HamiltonianProposalCycle
                subclass); device proposals only
    prior_sampling : produce nlive samples from the prior and stop

    Device-only keywords: device (CUDA ordinal), batch (chains per kernel launch, default
    max(nthreads, nlive // 4)), lanes (lanes per chain, 0 = automatic), rho_target (chain-length
    control), check_functor (compare the device functor with the Python callables before running).
    """

    def __init__(self, usermodel, nlive=100, poolsize=100, output="./", verbose=0, seed=None, maxmcmc=100,
                 nthreads=None, nhamiltonian=0, nslice=0, resume=False, proposals=None, n_periodic_checkpoint=None,
                 periodic_checkpoint_interval=None, prior_sampling=False, device=0, batch=None, lanes=0,
                 rho_target=0.0, check_functor=True):
        self.nthreads = nthreads if nthreads is not None else (os.cpu_count() or 1)
        output = os.path.join(output, "")
        os.makedirs(output, exist_ok=True)
        self.verbose = verbose
        self.output = output
        self.logger = logging.getLogger("cpnest.cpnest.CPNest")
        self.log_file = LogFile(os.path.join(self.output, "cpnest.log"), verbose=self.verbose)
        with self.log_file:
            n_mh = self.nthreads - nhamiltonian - nslice
            if n_mh < 0 or nslice < 0 or nhamiltonian < 0:
                raise ValueError("nthreads={0} < nslice={1} + nhamiltonian={2}".format(self.nthreads, nslice, nhamiltonian))
            self.population = dict(mh=n_mh, slice=nslice, hmc=nhamiltonian)
            kinds = [k for k in ("mh", "slice", "hmc") if self.population[k] > 0]
            if not kinds:
                raise ValueError("no sampler requested: nthreads={0}".format(self.nthreads))
            self.sampler_kind = kinds[0] if len(kinds) == 1 else "+".join(kinds)
            primary = kinds[0]
            if n_periodic_checkpoint is not None:
                self.logger.critical("The n_periodic_checkpoint kwarg is deprecated, use periodic_checkpoint_interval instead.")
            from .proposal import DefaultProposalCycle, EnsembleSliceProposalCycle, HamiltonianProposalCycle
            defaults = dict(mhs=DefaultProposalCycle, sli=EnsembleSliceProposalCycle, hmc=HamiltonianProposalCycle)   # cpnest/cpnest.py:155-158
            if proposals is None:
                proposals = defaults
            elif isinstance(proposals, list):
                proposals = dict(mhs=proposals[0])
            proposals = dict(defaults, **proposals)
            self.nlive = nlive
            self.poolsize = poolsize
            self.posterior_samples = None
            self.logZ_error = None
            self.nested_samples = None
            self.prior_samples = None
            self.mcmc_samples = None
            self.prior_sampling = prior_sampling
            self.user = usermodel
            self.resume = resume
            # the reference pickles NestedSampler and every sampler (nested_sampler_resume.pkl, sampler_{i}.pkl,
            # cpnest/cpnest.py:176-179, 221-222); the device run is ONE state blob, kept under the NestedSampler's file name
            self.resume_file = os.path.join(output, "nested_sampler_resume.pkl")
            self.periodic_checkpoint_interval = (np.inf if periodic_checkpoint_interval is None
                                                 else float(periodic_checkpoint_interval))
            self.seed = 1234 if seed is None else seed
            self.maxmcmc = maxmcmc
            self.batch = int(batch) if batch else max(min(self.nthreads, nlive // 2), nlive // 4, 1)
            self.batch = max(len(kinds), min(self.batch, nlive - 3))
            self.logger.critical("Running on CUDA device {0} with {1} chains per batch".format(device, self.batch))
            # NestedSampler seeds numpy, then the proposal cycle is drawn (cpnest/cpnest.py:184-207)
            self.NS = NestedSampler(self.user, nlive=nlive, output=output, verbose=verbose, seed=self.seed,
                                    prior_sampling=self.prior_sampling)
            functor, params = self.user.get_device_functor()
            self.engine = Engine(functor, self.user.bounds, nlive=nlive, maxmcmc=maxmcmc, seed=self.seed, params=params,
                                 device=device, lanes=lanes, rho_target=rho_target, sampler=primary,
                                 mix=self.population if len(kinds) > 1 else None)
            # verbose >= 3: every sampler process of the reference keeps the last 5*maxmcmc states of its chains and saves
            # them as mcmc_chain_<pid>.dat (cpnest/sampler.py:86, 261-267); here the MH kernel logs the first chains of every
            # batch (slot i plays sampler process i)
            self.chain_log_slots = 0
            if verbose >= 3 and not self.prior_sampling and "mh" in kinds:
                self.chain_log_slots = min(4, self.batch, max(1, self.nthreads))
                self.engine.enable_chain_log(self.chain_log_slots, 3 * 5 * maxmcmc + 64)
            # one proposal object per sampler kind present, in the reference's order of construction
            # (cpnest/cpnest.py:197-261: MH samplers, then HMC, then slice); self.proposal is the first kind's
            self.proposals = {}
            if "mh" in kinds:
                self.proposals["mh"] = proposals["mhs"]()                       # cpnest/cpnest.py:207
                self.engine.set_cycle(self.proposals["mh"].device_cycle())
            if "hmc" in kinds:
                self.proposals["hmc"] = proposals["hmc"](model=self.user)       # cpnest/cpnest.py:231
                self.proposals["hmc"].device_cycle()
                if functor in ("gaussian_mean_sigma", "gaussian_mixture"):
                    self.logger.warning("HMC with a non-flat prior inherits the reference's trajectory-draw + energy-test rule, "
                                        "which is not exactly reversible when the potential varies (DESIGN.md); prefer MH or slice")
            if "slice" in kinds:
                self.proposals["slice"] = proposals["sli"](model=self.user)     # cpnest/cpnest.py:252
                self.engine.set_slice_cycle(self.proposals["slice"].device_cycle())
            self.proposal = self.proposals[primary]
            if check_functor:
                self._check_functor()

    # ------------------------------------------------------------------------------------------
    def _check_functor(self, n=64, rtol=1e-9):
        """The device functor must agree with the model's own Python callables."""
        rs = np.random.RandomState(self.seed)
        b = np.asarray(self.user.bounds, dtype=np.float64)
        X = b[:, 0] + (b[:, 1] - b[:, 0]) * rs.uniform(size=(n, len(self.user.names)))
        logL, logP = self.engine.eval_model(X)
        for i in range(n):
            p = LivePoint(self.user.names, d=X[i].tolist())
            lp = float(self.user.log_prior(p))
            ll = float(np.asarray(self.user.log_likelihood(p)).reshape(-1)[0])
            for name, dev, host in (("log_prior", logP[i], lp), ("log_likelihood", logL[i], ll)):
                if np.isfinite(host) or np.isfinite(dev):
                    if not abs(dev - host) <= rtol * max(1.0, abs(host)):
                        raise ValueError("device functor %r disagrees with %s.%s at %s: device %r, python %r" % (
                            self.user.get_device_functor()[0], type(self.user).__name__, name, X[i].tolist(), dev, host))

    def run(self):
        """Run the sampler (cpnest/cpnest.py:263-316)."""
        with self.log_file:
            t0 = time.time()
            eng, NS, K = self.engine, self.NS, self.batch
            if self.resume and os.path.exists(self.resume_file):
                self._resume()                                  # cpnest/cpnest.py:181-183
            else:
                eng.init_prior()
                if self.verbose >= 3 and not self.prior_sampling:
                    # cpnest/sampler.py:137-152: at verbose >= 3 every sampler saves samples of the prior,
                    # prior_samples_<pid>.dat; here the initial live set IS nlive draws from the prior
                    ps = LivePointList(self.user.names, eng.live()).asnparray()
                    fn = "prior_samples_%d.dat" % os.getpid()
                    np.savetxt(os.path.join(self.output, fn), ps.ravel(), header=" ".join(ps.dtype.names), newline="\n", delimiter=" ")
                    self.logger.critical("Sampler process {0!s}: saved {1:d} prior samples in {2!s}".format(os.getpid(), len(ps), fn))
            NS.initialised = True
            if self.resume:                                     # cpnest/cpnest.py:271-277
                for sig in (signal.SIGTERM, signal.SIGALRM, signal.SIGQUIT, signal.SIGINT, signal.SIGUSR1, signal.SIGUSR2):
                    signal.signal(sig, sighandler)
            last_checkpoint = time.time()
            D = len(self.user.names)
            if self.prior_sampling:
                # NestedSampling.py:439-448: the live points ARE the prior samples
                NS.nested_samples = LivePointList(self.user.names, eng.live())
                NS.update_from(eng.state())
                NS.write_chain_to_file()
                NS.write_evidence_to_file()
                self.nested_samples = self.get_nested_samples(filename=None)
                self.prior_samples = self.get_prior_samples(filename=None)
                return
            # a signal (CheckPoint, raised by sighandler) may arrive anywhere in the loop body, also inside the periodic
            # checkpoint or while the state is read back: the whole loop and the final sweep sit inside the handler, which
            # writes the resume file from the device state as it is and exits with 130 (cpnest/cpnest.py:286-288)
            try:
                while True:
                    eng.run(K, max_batches=max(1, self.nlive // K), tolerance=NS.tolerance)
                    s = eng.state()
                    if time.time() - last_checkpoint > self.periodic_checkpoint_interval:      # NestedSampling.py:453-455
                        self.checkpoint()
                        last_checkpoint = time.time()
                    NS.update_from(s)
                    if self.verbose:
                        acc = s["total_accepted"] / max(1, s["total_steps"])
                        self.logger.info("{0:d}: n:{1:4d} acc:{2:.3f} H: {3:.2f} logLmin {4:.5f} dZ: {5:.3f} logZ: {6:.3f} logLmax: {7:.2f}".format(
                            s["iteration"], s["nmcmc"], acc, s["info"], s["logLmin"], s["condition"], s["logZ"], s["logLmax"]))
                    if not (s["condition"] > NS.tolerance) or s["done"]:
                        break
                NS.insertion_indices = eng.insertion_indices()
            except CheckPoint:
                self.checkpoint()
                sys.exit(130)
            # from here on the live set is consumed (final sweep): a signal can no longer be answered with a checkpoint
            if self.resume:
                for sig in (signal.SIGTERM, signal.SIGALRM, signal.SIGQUIT, signal.SIGINT, signal.SIGUSR1, signal.SIGUSR2):
                    signal.signal(sig, signal.SIG_DFL)
            rect, trap = eng.finalise()
            s = eng.state()
            NS.update_from(s)
            NS.logZ = trap
            NS.state.logZ = trap
            rows, lv = eng.dead(log_vols=True)
            NS.nested_samples = LivePointList(self.user.names, rows)
            NS.state.logLs = np.concatenate(([-np.inf], rows[:, D]))
            NS.state.log_vols = np.concatenate(([0.0], lv))
            NS.acceptance = s["total_accepted"] / max(1, s["total_steps"])
            self.total_evals = s["total_evals"]
            self.total_steps = s["total_steps"]
            NS.write_chain_to_file()
            NS.write_evidence_to_file()
            if os.path.exists(self.resume_file):                 # a finished run leaves no resume file (NestedSampling.py:490-492)
                os.remove(self.resume_file)
            self.logger.critical("Final evidence: {0:0.2f}".format(NS.state.logZ))
            self.logger.critical("Information: {0:.2f}".format(NS.state.info))
            # sqrt(H / nlive) is the error of a run that replaces one point per iteration; a batch removes its K worst points
            # before any replacement arrives, which inflates the variance of the log-volume per unit of shrinkage by
            # (1/(1-f) - 1) / (-ln(1-f)), f = K / nlive (profiles/r01_logz_validation.txt)
            f = min(K / float(self.nlive), 0.999)
            infl = (1.0 / (1.0 - f) - 1.0) / -np.log1p(-f)
            self.logZ_error = float(np.sqrt(max(NS.state.info, 0.0) / self.nlive * infl))
            self.logger.critical("Estimated error of the evidence: {0:.3f} (sqrt(H/nlive) = {1:.3f}, batches of {2:d})".format(
                self.logZ_error, float(np.sqrt(max(NS.state.info, 0.0) / self.nlive)), K))
            self.logger.critical("Sampling time: {0:.2f} s, {1:d} likelihood evaluations".format(time.time() - t0, s["total_evals"]))
            if self.verbose >= 2:
                NS.check_insertion_indices(rolling=False, filename="insertion_indices.dat")
                self.logger.critical("Saving nested samples in {0}".format(self.output))
                self.nested_samples = self.get_nested_samples()
                self.logger.critical("Saving posterior samples in {0}".format(self.output))
                self.posterior_samples = self.get_posterior_samples()
            else:
                NS.check_insertion_indices(rolling=False, filename=None)
                self.nested_samples = self.get_nested_samples(filename=None)
                self.posterior_samples = self.get_posterior_samples(filename=None)
            if self.verbose >= 3:                                # cpnest/cpnest.py:309-312
                self._write_mcmc_chains()
                self.prior_samples = self.get_prior_samples(filename=None)
                self.mcmc_samples = self.get_mcmc_samples(filename=None)
            if self.verbose >= 2:
                try:
                    self.plot(corner=False)
                except ImportError as e:
                    self.logger.warning("plotting skipped: {0}".format(e))

    # ------------------------------------------------------------------------------------------
    def get_nested_samples(self, filename="nested_samples.dat"):
        """Structured array of the nested samples (cpnest/cpnest.py:318-341)."""
        self.nested_samples = self.NS.nested_samples.asnparray()
        if filename:
            np.savetxt(os.path.join(self.NS.output_folder, filename), self.nested_samples.ravel(),
                       header=" ".join(self.nested_samples.dtype.names), newline="\n", delimiter=" ")
        return self.nested_samples

    def get_posterior_samples(self, filename="posterior.dat"):
        """Posterior samples by rejection on the nested-sampling weights (cpnest/cpnest.py:343-406)."""
        from .nest2pos import draw_posterior_many
        nested_samples = self.get_nested_samples()      # writes nested_samples.dat, as the reference does
        # weights from the integrator's own volume history: a batch of K removals shrinks the volume point by point with
        # 1/(N-j), which the reference's reconstruction in compute_weights (one removal per iteration) does not describe
        lv = getattr(self.NS.state, "log_vols", None)
        lv = [np.asarray(lv)] if lv is not None and len(lv) == len(nested_samples) + 1 else None
        posterior_samples = np.array(draw_posterior_many([nested_samples], [self.nlive], verbose=self.verbose, log_vols=lv))
        if self.prior_samples is None:
            self.prior_samples = {n: None for n in self.user.names}
        self.mcmc_samples = {n: None for n in self.user.names}
        if filename:
            np.savetxt(os.path.join(self.NS.output_folder, filename), posterior_samples.ravel(),
                       header=" ".join(posterior_samples.dtype.names), newline="\n", delimiter=" ")
        return posterior_samples

    def get_prior_samples(self, filename="prior.dat"):
        """cpnest/cpnest.py:408-443: the prior_samples_*.dat files written at verbose >= 3 (read and removed, as the
        reference does), and with prior_sampling the nested samples."""
        from numpy.lib.recfunctions import stack_arrays
        parts = []
        for f in sorted(os.listdir(self.NS.output_folder)):
            if "prior_samples" in f:
                parts.append(np.atleast_1d(np.genfromtxt(os.path.join(self.NS.output_folder, f), names=True)))
                os.remove(os.path.join(self.NS.output_folder, f))
        if self.NS.prior_sampling:
            parts.append(self.get_nested_samples(filename=None))
        if not parts:
            self.logger.critical("ERROR, no prior samples found!")
            return None
        prior_samples = parts[0] if len(parts) == 1 else stack_arrays(parts, usemask=False)
        if filename:
            np.savetxt(os.path.join(self.NS.output_folder, filename), prior_samples.ravel(),
                       header=" ".join(prior_samples.dtype.names), newline="\n", delimiter=" ")
        return prior_samples

    def _write_mcmc_chains(self):
        """cpnest/sampler.py:261-267: at verbose >= 3 every sampler saves its last 5*maxmcmc MCMC states as
        mcmc_chain_<pid>.dat; one file per logged chain slot here."""
        for slot in range(getattr(self, "chain_log_slots", 0)):
            hist = self.engine.chain_history(slot, last=5 * self.maxmcmc)
            if len(hist) == 0:
                continue
            arr = LivePointList(self.user.names, hist).asnparray()
            fn = "mcmc_chain_{0}_{1}.dat".format(os.getpid(), slot)
            np.savetxt(os.path.join(self.output, fn), arr.ravel(), header=" ".join(arr.dtype.names), newline="\n", delimiter=" ")
            self.logger.critical("Sampler process {0!s}: saved {1:d} mcmc samples in {2!s}".format(os.getpid(), len(arr), fn))

    def get_mcmc_samples(self, filename="mcmc.dat"):
        """Resampled MCMC samples (cpnest/cpnest.py:445-478): every mcmc_chain_*.dat of the output folder is thinned by its
        autocorrelation length and redrawn against the Metropolis-Hastings rule (nest2pos.resample_mcmc_chain), the files
        are removed, the chains stacked."""
        from numpy.lib.recfunctions import stack_arrays
        from .nest2pos import resample_mcmc_chain
        mcmc_samples = []
        for f in sorted(os.listdir(self.NS.output_folder)):
            if "mcmc_chain" in f:
                path = os.path.join(self.NS.output_folder, f)
                mcmc_samples.append(resample_mcmc_chain(np.atleast_1d(np.genfromtxt(path, names=True))))
                os.remove(path)
        if not mcmc_samples:
            self.logger.critical("ERROR, no MCMC samples found!")
            return None
        mcmc_samples = mcmc_samples[0] if len(mcmc_samples) == 1 else stack_arrays(mcmc_samples, usemask=False)
        if filename:
            np.savetxt(os.path.join(self.NS.output_folder, filename), mcmc_samples.ravel(),
                       header=" ".join(mcmc_samples.dtype.names), newline="\n", delimiter=" ")
        return mcmc_samples

    def plot(self, corner=True):
        """Diagnostic plots (cpnest/cpnest.py:480-531); needs matplotlib (+ corner)."""
        from . import plot
        pos = self.posterior_samples
        for n in pos.dtype.names:
            plot.plot_hist(pos[n].ravel(), name=n, filename=os.path.join(self.output, "posterior_{0}.pdf".format(n)))
        for n in self.nested_samples.dtype.names:
            plot.plot_chain(self.nested_samples[n], name=n, filename=os.path.join(self.output, "nschain_{0}.pdf".format(n)))
        plot.plot_indices(self.NS.insertion_indices, filename=os.path.join(self.output, "insertion_indices.pdf"))
        if corner:
            arr = np.stack([pos[n] for n in pos.dtype.names], axis=1)
            plot.plot_corner(arr, labels=pos.dtype.names, filename=os.path.join(self.output, "corner.pdf"))

    def checkpoint(self):
        """Write the run state to the resume file (cpnest/cpnest.py:548-550, NestedSampling.py:495-501)."""
        self.logger.critical("Checkpointing nested sampling")
        rec = dict(blob=self.engine.save_checkpoint(), names=list(self.user.names), nlive=self.nlive, seed=self.seed,
                   batch=self.batch, sampler=self.sampler_kind, maxmcmc=self.maxmcmc)
        tmp = self.resume_file + ".tmp"
        with open(tmp, "wb") as f:
            pickle.dump(rec, f)
        os.replace(tmp, self.resume_file)

    def _resume(self):
        with open(self.resume_file, "rb") as f:
            rec = pickle.load(f)
        for k, mine in (("names", list(self.user.names)), ("nlive", self.nlive), ("batch", self.batch),
                        ("sampler", self.sampler_kind), ("maxmcmc", self.maxmcmc)):
            if rec[k] != mine:
                raise ValueError("resume file %s belongs to a different run: %s = %r, this run has %r" % (
                    self.resume_file, k, rec[k], mine))
        self.engine.load_checkpoint(rec["blob"])
        self.logger.critical("Resuming NestedSampler from " + self.resume_file)
