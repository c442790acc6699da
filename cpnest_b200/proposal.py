"""Proposal descriptors.  In the reference these classes DO the jump in Python
(cpnest/proposal.py:14-362); here the arithmetic lives in the CUDA chain kernel
(csrc/mh_kernel.cuh) and the classes only select it: a ProposalCycle is drawn on the host exactly
like the reference's (np.random.choice over the normalised weights, proposal.py:71-75) and handed
to the device as one proposal id per cycle slot.
"""
import numpy as np

from ._lib import DE, EIGEN, SLICE_CORR, SLICE_DIFF, SLICE_GAUSS, STRETCH, WALK


class Proposal(object):
    """Base class (cpnest/proposal.py:14-34).  A proposal runs on the device iff it has `device_id`."""
    log_J = 0.0
    device_id = None

    def get_sample(self, old):
        raise NotImplementedError(
            "%s.get_sample: proposals are executed inside the CUDA chain kernel; Python-side jump proposals "
            "cannot run on the device and cpnest_b200 has no CPU fallback" % type(self).__name__)


class EnsembleProposal(Proposal):
    ensemble = None

    def set_ensemble(self, ensemble):
        self.ensemble = ensemble


class EnsembleWalk(EnsembleProposal):
    """Goodman & Weare walk move over 3 ensemble members (cpnest/proposal.py:209-235)."""
    device_id = WALK
    Npoints = 3


class EnsembleStretch(EnsembleProposal):
    """Goodman & Weare stretch move, scale 2, log_J = D * x (cpnest/proposal.py:237-261)."""
    device_id = STRETCH


class DifferentialEvolution(EnsembleProposal):
    """old + (b - a) * N(1, 1e-4) (cpnest/proposal.py:263-287)."""
    device_id = DE


class EnsembleEigenVector(EnsembleProposal):
    """Jump along a random eigenvector of the ensemble covariance (cpnest/proposal.py:289-342);
    covariance and eigen-decomposition are computed on the device (csrc/ensemble.cuh)."""
    device_id = EIGEN


class ProposalCycle(EnsembleProposal):
    """cpnest/proposal.py:47-112"""
    idx = 0
    N = 0

    def __init__(self, proposals, weights, cyclelength=100, *args, **kwargs):
        assert len(weights) == len(proposals)
        self.cyclelength = cyclelength
        self.weights = weights
        self.proposals = proposals
        self.set_cycle()

    @property
    def weights(self):
        return self._weights

    @weights.setter
    def weights(self, weights):
        norm = float(sum(weights))
        self._weights = [w / norm for w in weights]

    def set_cycle(self):
        # same draw as the reference: indices are chosen with the legacy global numpy RandomState
        idx = np.random.choice(len(self.proposals), size=self.cyclelength, p=self.weights, replace=True)
        self.cycle = [self.proposals[i] for i in idx]
        self.N = len(self.cycle)

    def add_proposal(self, proposal, weight):
        self.proposals = self.proposals + [proposal]
        self.weights = self.weights + [weight]
        self.set_cycle()

    def set_ensemble(self, ensemble):
        self.ensemble = ensemble

    def device_cycle(self):
        """One device proposal id per cycle slot."""
        ids = []
        for p in self.cycle:
            if getattr(p, "device_id", None) is None:
                raise NotImplementedError(
                    "proposal %s has no device implementation; cpnest_b200 runs EnsembleWalk, EnsembleStretch, "
                    "DifferentialEvolution and EnsembleEigenVector on the GPU and has no CPU fallback" % type(p).__name__)
            ids.append(p.device_id)
        return np.array(ids, dtype=np.uint8)


class DefaultProposalCycle(ProposalCycle):
    """Walk : Stretch : DE : EigenVector = 5 : 5 : 10 : 10 (cpnest/proposal.py:345-362)."""

    def __init__(self):
        proposals = [EnsembleWalk(), EnsembleStretch(), DifferentialEvolution(), EnsembleEigenVector()]
        super(DefaultProposalCycle, self).__init__(proposals, [5, 5, 10, 10])


class EnsembleSlice(EnsembleProposal):
    """Direction proposals of the ensemble slice sampler (cpnest/proposal.py:114-119); the directions are
    drawn inside the CUDA slice kernel (csrc/slice_kernel.cuh)."""
    log_J = 0.0

    def get_direction(self, mu=1.0):
        raise NotImplementedError("%s.get_direction: slice directions are drawn inside the CUDA slice kernel" % type(self).__name__)


class EnsembleSliceDifferential(EnsembleSlice):
    """Difference of two ensemble members times mu (cpnest/proposal.py:121-133)."""
    device_id = SLICE_DIFF


class EnsembleSliceGaussian(EnsembleSlice):
    """Isotropic unit direction times mu (cpnest/proposal.py:175-187)."""
    device_id = SLICE_GAUSS


class EnsembleSliceCorrelatedGaussian(EnsembleSlice):
    """mu * N(mean, cov) of the ensemble (cpnest/proposal.py:135-173)."""
    device_id = SLICE_CORR


class EnsembleSliceProposalCycle(ProposalCycle):
    """Differential : Gaussian : CorrelatedGaussian = 1 : 1 : 1 (cpnest/proposal.py:189-207)."""

    def __init__(self, model=None):
        proposals = [EnsembleSliceDifferential(), EnsembleSliceGaussian(), EnsembleSliceCorrelatedGaussian()]
        super(EnsembleSliceProposalCycle, self).__init__(proposals, [1, 1, 1])

    def get_direction(self, mu=1.0, **kwargs):
        raise NotImplementedError("slice directions are drawn inside the CUDA slice kernel")

    def device_cycle(self):
        ids = []
        for p in self.cycle:
            if not isinstance(p, EnsembleSlice) or getattr(p, "device_id", None) is None:
                raise NotImplementedError("slice direction %s has no device implementation" % type(p).__name__)
            ids.append(p.device_id)
        return np.array(ids, dtype=np.uint8)


class HamiltonianProposal(EnsembleEigenVector):
    """Base of the Hamiltonian proposals (cpnest/proposal.py:377-588): mass matrix = inverse ensemble covariance,
    dt / L adaptation.  Runs inside the CUDA HMC kernel (csrc/hmc_kernel.cuh)."""
    device_id = None
    TARGET = 0.500

    def __init__(self, model=None, **kwargs):
        self.model = model


class LeapFrog(HamiltonianProposal):
    """Unconstrained leapfrog (cpnest/proposal.py:590-673); on the device it is the constrained kernel at logLmin = -inf."""
    hmc_device = True


class ConstrainedLeapFrog(LeapFrog):
    """Leapfrog with reflection off the logL = logLmin surface (cpnest/proposal.py:675-844)."""
    hmc_device = True


class HamiltonianProposalCycle(ProposalCycle):
    """One ConstrainedLeapFrog (cpnest/proposal.py:364-375)."""

    def __init__(self, model=None):
        super(HamiltonianProposalCycle, self).__init__([ConstrainedLeapFrog(model=model)], [1])

    def device_cycle(self):
        for p in self.cycle:
            if not getattr(p, "hmc_device", False):
                raise NotImplementedError("Hamiltonian proposal %s has no device implementation" % type(p).__name__)
        return np.zeros(len(self.cycle), dtype=np.uint8)
