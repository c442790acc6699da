"""Host view of the nested-sampling driver.

The reference's NestedSampler (cpnest/NestedSampling.py:176-537) runs the loop in Python and talks
to sampler processes over pipes.  Here the loop runs on the device (csrc/ns_kernels.cuh) and this
class only mirrors the observable surface: `logZ`, `state.info`, `Nlive`, `iteration`,
`insertion_indices`, `nested_samples`, the output files and the insertion-index KS check.
"""
import logging
import os

import numpy as np
from scipy.stats import kstest

from .parameter import LivePoint


class _IntegralStateView(object):
    """Observable fields of _NSintegralState (cpnest/NestedSampling.py:88-144)."""

    def __init__(self, nlive):
        self.nlive = nlive
        self.iteration = 0
        self.logZ = -np.inf
        self.logw = 0.0
        self.info = 0.0
        self.logLs = [-np.inf]
        self.log_vols = [0.0]

    def plot(self, filename):
        import matplotlib as mpl
        mpl.use("Agg")
        from matplotlib import pyplot as plt
        plt.figure()
        plt.plot(self.log_vols, self.logLs)
        plt.title("{0} iterations. logZ={1:.2f} H={2:.2f} bits".format(self.iteration, self.logZ, self.info * np.log2(np.e)))
        plt.grid(which="both")
        plt.xlabel("log prior_volume")
        plt.ylabel("log likelihood")
        plt.xlim([self.log_vols[-1], self.log_vols[0]])
        plt.savefig(filename)


class LivePointList(object):
    """Sequence of LivePoint backed by an (n, D+2) record array: NS.nested_samples without
    materialising a Python object per dead point."""

    def __init__(self, names, rows):
        self.names = list(names)
        self.rows = rows

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return LivePointList(self.names, self.rows[i])
        r = self.rows[i]
        D = len(self.names)
        return LivePoint(self.names, d=r[:D].tolist(), logL=r[D], logP=r[D + 1])

    def __iter__(self):
        for i in range(len(self.rows)):
            yield self[i]

    def asnparray(self):
        names = self.names + ["logL", "logPrior"]
        out = np.zeros(len(self.rows), dtype={"names": names, "formats": ["f8"] * len(names)})
        for k, n in enumerate(names):
            out[n] = self.rows[:, k]
        return out


class NestedSampler(object):
    def __init__(self, model, nlive=1024, output=None, verbose=1, seed=1, prior_sampling=False, stopping=0.1):
        self.logger = logging.getLogger("cpnest.NestedSampling.NestedSampler")
        self.model = model
        self.prior_sampling = prior_sampling
        self.seed = seed
        np.random.seed(seed=self.seed)               # setup_random_seed (NestedSampling.py:311-316)
        self.verbose = verbose
        self.acceptance = 1.0
        self.Nlive = nlive
        self.tolerance = stopping
        self.condition = np.inf
        self.iteration = 0
        self.jumps = 0
        self.nested_samples = LivePointList(model.names, np.zeros((0, len(model.names) + 2)))
        self.insertion_indices = []
        self.rolling_p = []
        self.logZ = None
        self.logLmin = -np.inf
        self.logLmax = -np.inf
        self.state = _IntegralStateView(self.Nlive)
        self.output_folder = output
        chain = "chain_" + str(self.Nlive) + "_" + str(self.seed) + ".txt"
        self.output_file = os.path.join(output, chain)
        self.evidence_file = os.path.join(output, chain + "_evidence.txt")
        self.resume_file = os.path.join(output, "nested_sampler_resume.pkl")
        with open(os.path.join(output, "header.txt"), "w") as f:    # NestedSampling.py:264-267
            f.write("\t".join(self.model.names))
            f.write("\tlogL\n")
        self.initialised = False

    # ---- mirrors of the device state -------------------------------------------------------
    def update_from(self, s):
        self.state.logZ = s["logZ"]
        self.state.info = s["info"]
        self.state.logw = s["logw"]
        self.state.iteration = s["n_hist"]
        self.iteration = s["iteration"]
        self.condition = s["condition"]
        self.logLmin = s["logLmin"]
        self.logLmax = s["logLmax"]
        self.jumps = s["nmcmc"]

    # ---- output files ------------------------------------------------------------------------
    def write_chain_to_file(self):
        """chain_{nlive}_{seed}.txt (NestedSampling.py:293-301, model.py:108-127)."""
        rows = self.nested_samples.rows
        D = len(self.model.names)
        from .model import Model
        with open(self.output_file, "w") as f:
            f.write("{0:s}\n".format(self.model.header().rstrip()))
            if type(self.model).strsample is Model.strsample:
                # same bytes as '\t'.join('{:.20e}') + '{:20e}', formatted in C
                fmt = "\t".join(["%.20e"] * D) + "%20e"
                np.savetxt(f, rows[:, :D + 1], fmt=fmt, newline="\n")
            else:
                for ns in self.nested_samples:
                    f.write("{0:s}\n".format(self.model.strsample(ns).rstrip()))

    def write_evidence_to_file(self):
        """NestedSampling.py:303-309"""
        with open(self.evidence_file, "w") as f:
            f.write("#logZ\tlogLmax\tH\n")
            f.write("{0:.5f} {1:.5f} {2:.2f}\n".format(self.state.logZ, self.logLmax, self.state.info))

    def check_insertion_indices(self, rolling=True, filename=None):
        """KS test of the insertion indices against U(0,1) (NestedSampling.py:377-401)."""
        if len(self.insertion_indices) == 0:
            return
        indices = self.insertion_indices[-self.Nlive:] if rolling else self.insertion_indices
        D, p = kstest(indices, "uniform", args=(0, 1))
        if rolling:
            self.logger.warning("Rolling KS test: D={0:.3}, p-value={1:.3}".format(D, p))
            self.rolling_p.append(p)
        else:
            self.logger.warning("Final KS test: D={0:.3}, p-value={1:.3}".format(D, p))
        if filename is not None:
            np.savetxt(os.path.join(self.output_folder, filename), self.insertion_indices, newline="\n", delimiter=" ")
        return D, p
