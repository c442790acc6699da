"""Engine: numpy-facing wrapper of one device context (one GPU).

This is the object the `CPNest` facade drives; it owns what the reference spreads over
NestedSampler, its samplers and the RunManager (cpnest/cpnest.py:181-261): the live set, the
proposal cycle, the evidence integrator and the chain-length adaptation, all device resident.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from . import _lib as L_


def functor_blob(name, dim, params=None):
    """Pack the parameters of a device functor (include/cpnest_b200.h enum cpn_functor)."""
    params = dict(params or {})
    if name == "gaussian_iid":
        mu = float(params.get("mu", 0.0))
        sigma = float(params.get("sigma", 1.0))
        return np.array([mu, 1.0 / sigma, -math.log(sigma) - 0.5 * math.log(2.0 * math.pi)])
    if name == "gaussian_corr":
        Linv = np.asarray(params["Linv"], dtype=np.float64)
        assert Linv.shape == (dim, dim)
        Linv = np.tril(Linv)
        LinvT = np.zeros((dim, (dim + 31) // 32 * 32))       # transposed copy, rows padded with zeros to a multiple of 32
        LinvT[:, :dim] = Linv.T
        return np.concatenate(([float(params["logdet"])], Linv.ravel(), LinvT.ravel()))
    if name == "ring":
        return np.array([float(params.get("radius", 1.0)), float(params.get("thickness", 0.1))])
    if name == "gaussian_mixture":
        data = np.asarray(params["data"], dtype=np.float64).ravel()
        return np.concatenate(([float(len(data))], data))
    if name in ("eggbox", "rosenbrock", "ackley", "gaussian_mean_sigma"):
        return np.zeros(0)
    if name == "gaussian_mixture_nd":
        return np.array([float(params.get("w", 0.3)), float(params.get("m1", -2.5)), float(params.get("s1", 1.0)),
                         float(params.get("m2", 2.5)), float(params.get("s2", 1.0))])
    if name == "user":
        return np.asarray(params.get("params", []), dtype=np.float64).ravel()
    raise KeyError("unknown device functor %r (known: %s)" % (name, ", ".join(sorted(L.FUNCTORS))))


class Engine:
    def __init__(self, functor, bounds, nlive, maxmcmc=100, seed=1234, params=None, device=0,
                 nmcmc_init=0, lanes=0, ns_mode="sequential", nmcmc_safety=5.0, adapt_nmcmc=True,
                 rho_target=0.0, rank=0, world_size=1, sampler="mh", mix=None, two_lag=True, cold_eigen=False):
        self.lib = L.load()
        self.functor = functor
        self.bounds = np.asarray(bounds, dtype=np.float64).reshape(-1, 2)
        self.dim = self.bounds.shape[0]
        self.nlive = int(nlive)
        self.maxmcmc = int(maxmcmc)
        cfg = L.cpn_config()
        cfg.device = device
        cfg.dim = self.dim
        cfg.nlive = self.nlive
        cfg.maxmcmc = self.maxmcmc
        cfg.nmcmc_init = int(nmcmc_init)
        cfg.functor = L.FUNCTORS[functor]
        cfg.lanes = int(lanes)
        cfg.ns_mode = dict(sequential=L.NS_SEQUENTIAL, reference=L.NS_REFERENCE)[ns_mode]
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.nmcmc_safety = float(nmcmc_safety)
        cfg.adapt_nmcmc = 1 if adapt_nmcmc else 0
        cfg.rho_target = float(rho_target)
        cfg.flags = (0 if two_lag else 1) | (2 if cold_eigen else 0)      # CPN_NO_TWO_LAG, CPN_COLD_EIGEN
        cfg.world_size = int(world_size)
        cfg.rank = int(rank)
        if sampler not in L.SAMPLERS:
            raise ValueError("unknown sampler %r (known: %s)" % (sampler, ", ".join(L.SAMPLERS)))
        cfg.sampler = L.SAMPLERS[sampler]
        self.sampler = sampler
        # mixed population: mix = dict(mh=a, slice=b, hmc=c) or a 3-sequence in the order of enum cpn_sampler
        # (relative numbers of chains per batch, CPNest(nthreads, nslice, nhamiltonian), cpnest/cpnest.py:197-261)
        if mix is not None:
            if isinstance(mix, dict):
                unknown = set(mix) - set(L.SAMPLERS)
                if unknown:
                    raise ValueError("unknown sampler(s) in mix: %s" % ", ".join(sorted(unknown)))
                mix = [int(mix.get(k, 0)) for k in ("mh", "slice", "hmc")]
            mix = [int(m) for m in mix]
            if len(mix) != 3 or min(mix) < 0 or sum(mix) == 0:
                raise ValueError("mix must hold three non-negative numbers (mh, slice, hmc), not all zero")
            for i, m in enumerate(mix):
                cfg.mix[i] = m
        self.mix = mix
        blob = L.f64(functor_blob(functor, self.dim, params))
        lo = L.f64(self.bounds[:, 0])
        hi = L.f64(self.bounds[:, 1])
        h = C.c_void_p()
        L.check(self.lib.cpn_create(C.byref(cfg), L.dptr(lo), L.dptr(hi),
                                    blob.ctypes.data_as(C.c_void_p), blob.nbytes, C.byref(h)))
        self.h = h
        self.ns_mode = ns_mode
        if functor == "user":
            src = (params or {}).get("source")
            if not src:
                self.close()
                raise ValueError("functor 'user' needs params['source']: CUDA C++ defining cpn_user_log_likelihood "
                                 "(contract in cpnest_b200/csrc/user_model.cuh)")
            self.set_user_functor(src)

    # ---- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.cpn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- set-up -----------------------------------------------------------------------------
    def set_cycle(self, cycle):
        cyc = np.ascontiguousarray(cycle, dtype=np.uint8)
        L.check(self.lib.cpn_set_cycle(self.h, cyc.ctypes.data_as(C.POINTER(C.c_uint8)), len(cyc)))

    def set_slice_cycle(self, cycle):
        cyc = np.ascontiguousarray(cycle, dtype=np.uint8)
        L.check(self.lib.cpn_set_slice_cycle(self.h, cyc.ctypes.data_as(C.POINTER(C.c_uint8)), len(cyc)))

    def set_slice_mu(self, mu):
        L.check(self.lib.cpn_set_slice_mu(self.h, float(mu)))

    def set_user_functor(self, source):
        """Compile a user-written CUDA functor (NVRTC, sm_100a) into this context's chain kernels."""
        import os
        here = os.path.dirname(os.path.abspath(__file__))
        dirs = [os.path.join(here, "csrc"), os.path.join(os.path.dirname(here), "include")]
        arr = (C.c_char_p * len(dirs))(*[d.encode() for d in dirs])
        try:
            L.check(self.lib.cpn_set_user_functor(self.h, source.encode(), arr, len(dirs)))
        except Exception:
            self.close()
            raise

    def set_seed(self, seed):
        L.check(self.lib.cpn_set_seed(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def init_prior(self):
        L.check(self.lib.cpn_init_prior(self.h))

    def set_live(self, rows):
        rows = L.f64(rows)
        assert rows.shape == (self.nlive, self.dim + 2)
        L.check(self.lib.cpn_set_live(self.h, L.dptr(rows), self.nlive))

    def update_ensemble_stats(self):
        L.check(self.lib.cpn_update_ensemble_stats(self.h))

    def ensemble_stats(self):
        D = self.dim
        mean, cov, w, v = np.zeros(D), np.zeros((D, D)), np.zeros(D), np.zeros((D, D))
        L.check(self.lib.cpn_get_ensemble_stats(self.h, L.dptr(mean), L.dptr(cov), L.dptr(w), L.dptr(v)))
        return mean, cov, w, v

    # ---- the hot path -----------------------------------------------------------------------
    def select_and_accumulate(self, K):
        L.check(self.lib.cpn_select_and_accumulate(self.h, K))

    def evolve_batch(self, K):
        L.check(self.lib.cpn_evolve_batch(self.h, K))

    def insert_batch(self, K):
        L.check(self.lib.cpn_insert_batch(self.h, K))

    def ns_step(self, K):
        L.check(self.lib.cpn_ns_step(self.h, K))

    def run(self, K, max_batches=0, tolerance=0.1):
        done = C.c_int64(0)
        L.check(self.lib.cpn_run(self.h, K, max_batches, tolerance, C.byref(done)))
        return done.value

    def finalise(self):
        rect, trap = C.c_double(), C.c_double()
        L.check(self.lib.cpn_finalise(self.h, C.byref(rect), C.byref(trap)))
        return rect.value, trap.value

    def save_checkpoint(self):
        """The run state as bytes (cpn_save_checkpoint)."""
        n = self.lib.cpn_checkpoint_bytes(self.h)
        if n < 0:
            raise L.CpnError(self.lib.cpn_last_error().decode())
        buf = np.empty(n, dtype=np.uint8)
        L.check(self.lib.cpn_save_checkpoint(self.h, buf.ctypes.data_as(C.c_void_p), n))
        return buf.tobytes()

    def load_checkpoint(self, blob):
        buf = np.frombuffer(blob, dtype=np.uint8)
        L.check(self.lib.cpn_load_checkpoint(self.h, buf.ctypes.data_as(C.c_void_p), len(buf)))

    def synchronize(self):
        L.check(self.lib.cpn_synchronize(self.h))

    def set_stream(self, cuda_stream):
        L.check(self.lib.cpn_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def enable_chain_log(self, n_chains, n_events):
        """Record the history of the first n_chains chains of every MH batch (cpn_enable_chain_log)."""
        L.check(self.lib.cpn_enable_chain_log(self.h, int(n_chains), int(n_events)))
        self._chain_log_events = int(n_events)

    def chain_history(self, slot, last=None):
        """Per-step states of the chains that ran in `slot`, oldest first, as rows [x.., logL, logP] - the device form of
        the sampler's `samples` deque (cpnest/sampler.py:345: one entry per MCMC step).  Rebuilt from the event log: a
        state is repeated for every step up to the next accepted one.  `last`: keep only the most recent `last` states."""
        cap = self._chain_log_events
        ev = np.empty((cap, self.dim + 4))
        n = C.c_int32(0)
        L.check(self.lib.cpn_get_chain_log(self.h, int(slot), L.dptr(ev), cap, C.byref(n)))
        ev = ev[:n.value]
        # the ring may have cut the oldest chain: start at the first chain start
        starts = np.where((ev[:, 0] == 0) & (ev[:, 1] == -1))[0]
        if len(starts) == 0:
            return np.empty((0, self.dim + 2))
        ev = ev[starts[0]:]
        out = []
        i = 0
        while i < len(ev):
            j = i + 1
            while j < len(ev) and not (ev[j, 0] == 1):
                j += 1
            if j >= len(ev):
                break                                     # a chain without its end event: still running / cut
            steps = int(ev[j, 1])
            seg = ev[i:j]                                 # state events of this chain, first one at step -1
            at = seg[:, 1].astype(np.int64)
            idx = np.searchsorted(at, np.arange(steps), side="right") - 1     # state after step s: last event with step <= s
            rows = np.concatenate((seg[idx, 4:], seg[idx, 2:4]), axis=1)
            out.append(rows)
            i = j + 1
        if not out:
            return np.empty((0, self.dim + 2))
        hist = np.concatenate(out)
        return hist[-last:] if last else hist

    def pick_lanes(self, chains):
        """(lanes per chain, coordinates per lane) of a launch of `chains` chains (cpn_pick_lanes)."""
        g, e = C.c_int32(), C.c_int32()
        L.check(self.lib.cpn_pick_lanes(self.h, int(chains), C.byref(g), C.byref(e)))
        return g.value, e.value

    def launch_count(self):
        return self.lib.cpn_launch_count(self.h)

    def evolve_host(self, requests, logLmin, nmcmc, slots=None, replies=None, steps=None, accepted=None):
        """Sampler pipe protocol with host buffers (numpy arrays or pinned torch tensors'
        .numpy() views): requests[K][D+2] in, replies[K][D+2] out."""
        K = requests.shape[0]
        if replies is None:
            replies = np.empty_like(requests)
        if steps is None:
            steps = np.empty(K, dtype=np.int32)
        if accepted is None:
            accepted = np.empty(K, dtype=np.int32)
        L.check(self.lib.cpn_evolve_host(self.h, K, float(logLmin), int(nmcmc), requests.ctypes.data,
                                         slots.ctypes.data if slots is not None else None,
                                         replies.ctypes.data, steps.ctypes.data, accepted.ctypes.data))
        return replies, steps, accepted

    def evolve_host_table(self, table, start_idx, slots, logLmin, nmcmc, steps=None, accepted=None, rows_async=False):
        """The pipe protocol over the caller's pinned host live table (cpn_evolve_host_table): request j is row
        table[start_idx[j]], reply j lands in row table[slots[j]].  `table` is the .numpy() view of a pinned torch tensor
        (or any cudaHostRegister'ed C-contiguous float64 array) of shape [nlive][D+2].  rows_async: return once the
        logL / logPrior columns and steps / accepted are final; the coordinates of the replaced rows are complete after
        the next call on this engine or synchronize() (CPN_TABLE_ROWS_ASYNC)."""
        K = len(start_idx)
        assert table.flags.c_contiguous and table.dtype == np.float64 and table.shape == (self.nlive, self.dim + 2)
        start_idx = np.ascontiguousarray(start_idx, dtype=np.int32)
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        assert len(slots) == K
        if steps is None:
            steps = np.empty(K, dtype=np.int32)
        if accepted is None:
            accepted = np.empty(K, dtype=np.int32)
        L.check(self.lib.cpn_evolve_host_table(self.h, K, float(logLmin), int(nmcmc), table.ctypes.data, start_idx.ctypes.data,
                                               slots.ctypes.data, steps.ctypes.data, accepted.ctypes.data, 1 if rows_async else 0))
        return steps, accepted

    @staticmethod
    def kind_ranges(mix, K):
        """Chain ranges [kb[s], kb[s+1]) of the sampler kinds (mh, slice, hmc) in a batch of K chains (cpn_kind_ranges)."""
        m = np.ascontiguousarray(mix, dtype=np.int32)
        assert m.shape == (3,)
        kb = np.zeros(4, dtype=np.int32)
        L.check(L.load().cpn_kind_ranges(L.iptr(m), int(K), L.iptr(kb)))
        return kb.tolist()

    @staticmethod
    def host_gather_rows(table, idx, out):
        """out[j] = table[idx[j]] for C-contiguous float64 tables (cpn_host_gather_rows)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        assert table.flags.c_contiguous and out.flags.c_contiguous and table.dtype == out.dtype == np.float64
        assert out.shape == (len(idx), table.shape[1])
        L.check(L.load().cpn_host_gather_rows(table.ctypes.data, table.shape[0], table.shape[1], idx.ctypes.data, len(idx),
                                              out.ctypes.data))
        return out

    @staticmethod
    def host_scatter_rows(table, idx, rows):
        """table[idx[j]] = rows[j] (cpn_host_scatter_rows)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        assert table.flags.c_contiguous and rows.flags.c_contiguous and table.dtype == rows.dtype == np.float64
        assert rows.shape == (len(idx), table.shape[1])
        L.check(L.load().cpn_host_scatter_rows(table.ctypes.data, table.shape[0], table.shape[1], idx.ctypes.data, len(idx),
                                               rows.ctypes.data))

    # ---- observation ------------------------------------------------------------------------
    def state(self):
        s = L.cpn_state()
        L.check(self.lib.cpn_get_state(self.h, C.byref(s)))
        return s.as_dict()

    def set_nmcmc(self, n):
        L.check(self.lib.cpn_set_nmcmc(self.h, int(n)))

    def live(self):
        rows = np.empty((self.nlive, self.dim + 2))
        L.check(self.lib.cpn_get_live(self.h, L.dptr(rows)))
        return rows

    def num_dead(self):
        n = self.lib.cpn_num_dead(self.h)
        if n < 0:
            raise L.CpnError(self.lib.cpn_last_error().decode())
        return n

    def dead(self, first=0, n=None, log_vols=False):
        if n is None:
            n = self.num_dead() - first
        rows = np.empty((n, self.dim + 2))
        lv = np.empty(n) if log_vols else None
        L.check(self.lib.cpn_get_dead(self.h, first, n, L.dptr(rows), L.dptr(lv) if log_vols else None))
        return (rows, lv) if log_vols else rows

    def insertion_indices(self):
        n = self.lib.cpn_num_insertion_indices(self.h)
        out = np.empty(max(n, 0))
        if n > 0:
            L.check(self.lib.cpn_get_insertion_indices(self.h, 0, n, L.dptr(out)))
        return out

    def last_batch(self, K):
        rows = np.empty((K, self.dim + 2))
        steps = np.empty(K, dtype=np.int32)
        acc = np.empty(K, dtype=np.int32)
        L.check(self.lib.cpn_get_last_batch(self.h, K, L.dptr(rows), L.iptr(steps), L.iptr(acc)))
        return rows, steps, acc

    # ---- parity entry points ----------------------------------------------------------------
    def eval_model(self, X):
        X = L.f64(X).reshape(-1, self.dim)
        n = X.shape[0]
        logL, logP = np.empty(n), np.empty(n)
        L.check(self.lib.cpn_eval_model(self.h, L.dptr(X), n, L.dptr(logL), L.dptr(logP)))
        return logL, logP

    def eval_gradients(self, X):
        X = L.f64(X).reshape(-1, self.dim)
        n = X.shape[0]
        gl, gv = np.empty((n, self.dim)), np.empty((n, self.dim))
        L.check(self.lib.cpn_eval_gradients(self.h, L.dptr(X), n, L.dptr(gl), L.dptr(gv)))
        return gl, gv

    def replay_mh(self, ensemble, start, eigval, eigvec, cycle_idx, tape, nmcmc, maxmcmc, logLmin,
                  compat=True, lanes=0, record=False):
        ens = L.f64(ensemble)
        start = L.f64(start).reshape(-1, self.dim + 2)
        Cn = start.shape[0]
        tape = L.f64(tape).reshape(Cn, maxmcmc, 8)
        ev = L.f64(eigval)
        evec = L.f64(eigvec)
        out = np.empty((Cn, self.dim + 2))
        steps = np.empty(Cn, dtype=np.int32)
        acc = np.empty(Cn, dtype=np.int32)
        states = np.zeros((Cn, maxmcmc, self.dim)) if record else None
        L.check(self.lib.cpn_replay_mh(self.h, ens.shape[0], L.dptr(ens), Cn, L.dptr(start), L.dptr(ev), L.dptr(evec),
                                       int(cycle_idx), L.dptr(tape), int(nmcmc), int(maxmcmc), float(logLmin),
                                       1 if compat else 0, int(lanes), L.dptr(out), L.iptr(steps), L.iptr(acc),
                                       L.dptr(states) if record else None))
        return out, steps, acc, states

    def trace_mh(self, ensemble, start, eigval, eigvec, cycle_idx, nmcmc, maxmcmc, logLmin, lanes=0, chain_base=0):
        """The production MH kernel on given chains + the record of what every step consumed (cpn_trace_mh).
        Returns (out[C][D+2], steps, accepted, evals, draws[C][maxmcmc][8])."""
        ens = L.f64(ensemble)
        start = L.f64(start).reshape(-1, self.dim + 2)
        Cn = start.shape[0]
        ev = L.f64(eigval)
        evec = L.f64(eigvec)
        out = np.empty((Cn, self.dim + 2))
        steps = np.empty(Cn, dtype=np.int32)
        acc = np.empty(Cn, dtype=np.int32)
        evals = np.empty(Cn, dtype=np.int32)
        draws = np.zeros((Cn, maxmcmc, 8))
        L.check(self.lib.cpn_trace_mh(self.h, ens.shape[0], L.dptr(ens), Cn, L.dptr(start), L.dptr(ev), L.dptr(evec),
                                      int(cycle_idx), int(nmcmc), int(maxmcmc), float(logLmin), int(lanes), int(chain_base),
                                      L.dptr(out), L.iptr(steps), L.iptr(acc), L.iptr(evals), L.dptr(draws)))
        return out, steps, acc, evals, draws

    def replay_slice(self, ensemble, start, cycle_idx, tapes, mu, nmcmc, maxmcmc, logLmin, compat=True, lanes=0):
        """Slice chains from explicit random tapes (one flat float sequence per chain, see
        include/cpnest_b200.h cpn_replay_slice).  Returns (out, steps, accepted, evals, last)."""
        ens = L.f64(ensemble)
        start = L.f64(start).reshape(-1, self.dim + 2)
        Cn = start.shape[0]
        assert len(tapes) == Cn
        off = np.zeros(Cn + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in tapes])
        tape = L.f64(np.concatenate([np.asarray(t, dtype=np.float64) for t in tapes]))
        out = np.empty((Cn, self.dim + 2))
        steps = np.empty(Cn, dtype=np.int32)
        acc = np.empty(Cn, dtype=np.int32)
        evals = np.empty(Cn, dtype=np.int32)
        last = np.empty((Cn, 4))
        L.check(self.lib.cpn_replay_slice(self.h, ens.shape[0], L.dptr(ens), Cn, L.dptr(start), int(cycle_idx), L.dptr(tape),
                                          off.ctypes.data_as(C.POINTER(C.c_int64)), float(mu), int(nmcmc), int(maxmcmc),
                                          float(logLmin), 1 if compat else 0, int(lanes), L.dptr(out), L.iptr(steps),
                                          L.iptr(acc), L.iptr(evals), L.dptr(last)))
        return out, steps, acc, evals, last

    def trace_slice(self, ensemble, start, cycle_idx, mu, nmcmc, maxmcmc, logLmin, mean=None, eigval=None, eigvec=None, lanes=0,
                    chain_base=0, trace_cap=None):
        """The production slice kernel on given chains (include/cpnest_b200.h cpn_trace_slice).  Returns (out, steps, accepted,
        evals, last, tapes): tapes[c] is the list of events (kind, value) the chain consumed, in the vocabulary of the oracle's
        tape reader ("np_uniform", "sample", "np_normal", "np_mvn", "np_exponential"), kinds[c] the direction kind per slice."""
        ens = L.f64(ensemble)
        start = L.f64(start).reshape(-1, self.dim + 2)
        Cn, D = start.shape[0], self.dim
        cap = int(trace_cap or (maxmcmc + 2) * (D + 6 + 64))
        out = np.empty((Cn, D + 2))
        steps = np.empty(Cn, dtype=np.int32)
        acc = np.empty(Cn, dtype=np.int32)
        evals = np.empty(Cn, dtype=np.int32)
        last = np.empty((Cn, 4))
        trace = np.zeros((Cn, cap))
        stats = [None, None, None]
        if mean is not None:
            stats = [L.f64(mean), L.f64(eigval), L.f64(np.asarray(eigvec).reshape(D, D))]
        L.check(self.lib.cpn_trace_slice(self.h, ens.shape[0], L.dptr(ens), Cn, L.dptr(start),
                                         *[L.dptr(a) if a is not None else None for a in stats], int(cycle_idx), float(mu),
                                         int(nmcmc), int(maxmcmc), float(logLmin), int(lanes), int(chain_base), L.dptr(out),
                                         L.iptr(steps), L.iptr(acc), L.iptr(evals), L.dptr(last), L.dptr(trace), cap))
        tapes, kinds = [], []
        for c in range(Cn):
            t = trace[c]
            n = int(t[0])
            if n < 0:
                raise RuntimeError("trace_slice: trace_cap %d too small" % cap)
            ev, kd, i = [], [], 1
            while i < 1 + n:
                kind, nx = int(t[i]), int(t[i + 1])
                ev.append(("np_uniform", float(t[i + 2])))
                i += 3
                if kind == 0:
                    ev.append(("sample", (int(t[i]), int(t[i + 1]))))
                    i += 2
                else:
                    ev.append(("np_normal" if kind == 1 else "np_mvn", t[i:i + D].copy()))
                    i += D
                ev.append(("np_exponential", float(t[i])))
                ev.append(("np_uniform", float(t[i + 1])))
                i += 2
                ev.extend(("np_uniform", float(v)) for v in t[i:i + nx])
                i += nx
                kd.append(kind)
            tapes.append(ev)
            kinds.append(kd)
        return out, steps, acc, evals, last, tapes, kinds

    def trace_hmc(self, start, cov, eigval, eigvec, dt, nmcmc, max_traj, logLmin, lanes=0, chain_base=0, trace_cap=8192):
        """The production HMC kernel on given chains (include/cpnest_b200.h cpn_trace_hmc).  Returns (out, steps, accepted, evals,
        trajectories): trajectories[c] = list of dict(L, p0, u (the 2L-1 uniforms of the weighted draw), u_acc)."""
        start = L_.f64(start).reshape(-1, self.dim + 2)
        Cn, D = start.shape[0], self.dim
        cov = L_.f64(cov).reshape(D, D)
        eigval, eigvec = L_.f64(eigval), L_.f64(np.asarray(eigvec).reshape(D, D))
        out = np.empty((Cn, D + 2))
        steps = np.empty(Cn, dtype=np.int32)
        acc = np.empty(Cn, dtype=np.int32)
        evals = np.empty(Cn, dtype=np.int32)
        trace = np.zeros((Cn, int(trace_cap)))
        L_.check(self.lib.cpn_trace_hmc(self.h, Cn, L_.dptr(start), L_.dptr(cov), L_.dptr(eigval), L_.dptr(eigvec), float(dt), int(nmcmc),
                                        int(max_traj), float(logLmin), int(lanes), int(chain_base), L_.dptr(out), L_.iptr(steps),
                                        L_.iptr(acc), L_.iptr(evals), L_.dptr(trace), int(trace_cap)))
        trajs = []
        for c in range(Cn):
            t = trace[c]
            n = int(t[0])
            if n < 0:
                raise RuntimeError("trace_hmc: trace_cap %d too small" % trace_cap)
            lst, i = [], 1
            while i < 1 + n:
                Lt = int(t[i])
                lst.append(dict(L=Lt, p0=t[i + 1:i + 1 + D].copy(), u=t[i + 1 + D:i + D + 2 * Lt].copy(), u_acc=float(t[i + D + 2 * Lt])))
                i += D + 2 * Lt + 1
            trajs.append(lst)
        return out, steps, acc, evals, trajs

    def replay_hmc(self, start, cov, logdet, tapes, dt, L, logLmin, max_traj=1000, lanes=0):
        """HMC chains from explicit tapes (include/cpnest_b200.h cpn_replay_hmc).
        Returns (out, steps, accepted, evals, log_J)."""
        start = L_.f64(start).reshape(-1, self.dim + 2)
        Cn = start.shape[0]
        assert len(tapes) == Cn
        off = np.zeros(Cn + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in tapes])
        tape = L_.f64(np.concatenate([np.asarray(t, dtype=np.float64) for t in tapes]))
        cov = L_.f64(cov).reshape(self.dim, self.dim)
        out = np.empty((Cn, self.dim + 2))
        steps = np.empty(Cn, dtype=np.int32)
        acc = np.empty(Cn, dtype=np.int32)
        evals = np.empty(Cn, dtype=np.int32)
        logJ = np.empty(Cn)
        L_.check(self.lib.cpn_replay_hmc(self.h, Cn, L_.dptr(start), L_.dptr(cov), float(logdet), L_.dptr(tape),
                                         off.ctypes.data_as(C.POINTER(C.c_int64)), float(dt), int(L), int(max_traj),
                                         float(logLmin), int(lanes), L_.dptr(out), L_.iptr(steps), L_.iptr(acc),
                                         L_.iptr(evals), L_.dptr(logJ)))
        return out, steps, acc, evals, logJ

    def slice_directions(self, direction, n, mu=1.0):
        out = np.empty((n, self.dim))
        L.check(self.lib.cpn_slice_directions(self.h, int(direction), int(n), float(mu), L.dptr(out)))
        return out

    def stage_replacements(self, rows):
        rows = L.f64(rows).reshape(-1, self.dim + 2)
        L.check(self.lib.cpn_stage_replacements(self.h, rows.shape[0], L.dptr(rows)))

    def ns_reset(self):
        L.check(self.lib.cpn_ns_reset(self.h))

    def ns_increment(self, logL, nlive, mode="sequential"):
        logL = L.f64(np.atleast_1d(logL))
        m = dict(sequential=L.NS_SEQUENTIAL, reference=L.NS_REFERENCE)[mode]
        L.check(self.lib.cpn_ns_increment(self.h, L.dptr(logL), len(logL), int(nlive), m))

    def ns_trapezoid(self):
        z = C.c_double()
        L.check(self.lib.cpn_ns_trapezoid(self.h, C.byref(z)))
        return z.value

    def philox(self, ctr, key):
        ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
        key = np.ascontiguousarray(key, dtype=np.uint32).reshape(2)
        out = np.empty_like(ctr)
        u32p = C.POINTER(C.c_uint32)
        L.check(self.lib.cpn_philox(self.h, ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), ctr.shape[0],
                                    out.ctypes.data_as(u32p)))
        return out
