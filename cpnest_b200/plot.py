"""Diagnostic plots (cpnest/plot.py:1-121).  matplotlib / corner are optional and imported lazily:
they are not on the hot path and are absent from the build image."""


def _plt():
    import matplotlib as mpl
    mpl.use("Agg")
    from matplotlib import pyplot as plt
    return plt


def plot_chain(x, name=None, filename=None):
    plt = _plt()
    fig = plt.figure()
    plt.plot(x, ",")
    plt.grid()
    plt.xlabel("iteration")
    if name is not None:
        plt.ylabel(name)
    if filename is not None:
        plt.savefig(filename, bbox_inches="tight")
    plt.close(fig)


def plot_hist(x, name=None, prior_samples=None, mcmc_samples=None, filename=None):
    plt = _plt()
    fig = plt.figure()
    plt.hist(x, density=True, color="black", linewidth=1.25, histtype="step", bins=max(1, len(x) // 50), label="posterior")
    if prior_samples is not None:
        plt.hist(prior_samples, density=True, color="green", histtype="step", bins=max(1, len(x) // 50), label="prior")
    if mcmc_samples is not None:
        plt.hist(mcmc_samples, density=True, color="red", histtype="step", bins=max(1, len(x) // 50), label="mcmc")
    plt.legend(loc="upper left")
    plt.ylabel("probability density")
    if name is not None:
        plt.xlabel(name)
    if filename is not None:
        plt.savefig(filename, bbox_inches="tight")
    plt.close(fig)


def plot_indices(indices, filename=None, max_bins=30):
    import numpy as np
    plt = _plt()
    fig = plt.figure()
    plt.hist(indices, density=True, color="tab:blue", linewidth=1.25, histtype="step",
             bins=min(max(1, len(indices) // 100), max_bins))
    plt.axhline(1, color="black", linewidth=1.25, linestyle=":", label="pdf")
    plt.legend()
    plt.xlabel("insertion indices [0, 1]")
    if filename is not None:
        plt.savefig(filename, bbox_inches="tight")
    plt.close(fig)


def plot_corner(xs, ps=None, ms=None, filename=None, **kwargs):
    import corner
    fig = corner.corner(xs, color="k", show_titles=True, quantiles=[0.05, 0.5, 0.95], **kwargs)
    if filename is not None:
        fig.savefig(filename, bbox_inches="tight")
