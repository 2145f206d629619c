"""Small host-side helpers of the drop-in classes (no gym dependency)."""
import numpy as np


class LinearDecay(object):
    """epsilon / beta schedule: linear decay then flat (reference: DecayThenFlatSchedule(decay="linear"),
    utils/util.py:78-100)."""

    def __init__(self, start, finish, time_length):
        self.start, self.finish, self.time_length = start, finish, time_length
        self.delta = (start - finish) / time_length

    def eval(self, T):
        return max(self.finish, self.start - self.delta * T)


def space_dim(space):
    if isinstance(space, (list, tuple)):
        return int(space[0])
    name = space.__class__.__name__
    if name == "Box":
        return int(space.shape[0])
    if name == "Discrete":
        return int(space.n)
    raise NotImplementedError("B200 QMIX path supports Box / Discrete / [dim] spaces, got %r" % (space,))


def is_discrete(space):
    return space.__class__.__name__ == "Discrete" or "MultiDiscrete" in space.__class__.__name__


def onehot(idx, dim):
    return np.eye(dim, dtype=np.float64)[np.asarray(idx)]
