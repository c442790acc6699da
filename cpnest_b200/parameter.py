"""Host-side LivePoint: the value type of the Model contract.

On the device a live point is a row of the structure-of-arrays live set (x[N][Dp], logL[N],
logP[N]); this class is the host view handed to user callables (`Model.log_likelihood(p)`,
`p['name']`) and used for outputs.  Interface of cpnest/parameter.pyx:16-157: `names`, `values`
(array('d')), `logL`, `logP`, `dimension`, name-keyed item access, element-wise arithmetic,
`copy()`, `asnparray()`.  Like the reference, `*` and `/` take the scalar in single precision
(parameter.pyx:96-120 declare it as C `float`).
"""
from array import array

import numpy as np


def _f32(s):
    return float(np.float32(s))


class LivePoint(object):
    __slots__ = ("names", "values", "logL", "logP", "dimension", "_index")

    def __init__(self, names, d=None, logL=-np.inf, logP=-np.inf):
        self.names = names
        self.dimension = len(names)
        self.values = array("d", d) if d is not None else array("d", [0.0] * self.dimension)
        self.logL = float(logL)
        self.logP = float(logP)
        self._index = None

    def keys(self):
        return self.names

    def _pos(self, name):
        if self._index is None:
            self._index = {n: i for i, n in enumerate(self.names)}
        try:
            return self._index[name]
        except KeyError:
            raise KeyError(name)

    def __getitem__(self, name):
        return self.values[self._pos(name)]

    def __setitem__(self, name, value):
        self.values[self._pos(name)] = float(value)

    def __len__(self):
        return self.dimension

    def __repr__(self):
        return "LivePoint({0:s}, d={1:s}, logL={2:f}, logP={3:f})".format(
            str(self.names), str(self.values), self.logL, self.logP)

    def __str__(self):
        return str({n: self[n] for n in self.names})

    def __getstate__(self):
        return (self.logL, self.logP, self.values)

    def __setstate__(self, state):
        self.logL, self.logP = state[0], state[1]
        self.values = array("d", state[2])

    def __reduce__(self):
        return (LivePoint, (self.names,), self.__getstate__())

    def _new(self, vals):
        return LivePoint(self.names, d=vals)

    def __add__(self, other):
        assert self.dimension == other.dimension
        return self._new([a + b for a, b in zip(self.values, other.values)])

    def __iadd__(self, other):
        assert self.dimension == other.dimension
        for i in range(self.dimension):
            self.values[i] += other.values[i]
        return self

    def __sub__(self, other):
        assert self.dimension == other.dimension
        return self._new([a - b for a, b in zip(self.values, other.values)])

    def __isub__(self, other):
        assert self.dimension == other.dimension
        for i in range(self.dimension):
            self.values[i] -= other.values[i]
        return self

    def __mul__(self, other):
        s = _f32(other)
        return self._new([s * a for a in self.values])

    def __imul__(self, other):
        s = _f32(other)
        for i in range(self.dimension):
            self.values[i] *= s
        return self

    def __truediv__(self, other):
        s = _f32(other)
        return self._new([a / s for a in self.values])

    def __itruediv__(self, other):
        s = _f32(other)
        for i in range(self.dimension):
            self.values[i] /= s
        return self

    def copy(self):
        return LivePoint(self.names, d=self.values, logL=self.logL, logP=self.logP)

    def asnparray(self):
        """1-row structured array with the columns names + ['logL', 'logPrior'] (parameter.pyx:148-157)."""
        names = list(self.names) + ["logL", "logPrior"]
        x = np.zeros(1, dtype={"names": names, "formats": ["f8"] * len(names)})
        for n, v in zip(self.names, self.values):
            x[n] = v
        x["logL"] = self.logL
        x["logPrior"] = self.logP
        return x


def rebuild_livepoint(names):
    return LivePoint(names)
