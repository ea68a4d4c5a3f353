"""MultiDiscrete action space: a vector of independent discrete sub-actions, each with an
inclusive [min, max] range (reference: multiagent/multi_discrete.py:9-43).  Used when an agent
both moves and speaks (environment.py:58-61); `_set_action` splits the flat action vector by the
sub-space sizes `high - low + 1` (environment.py:148-155)."""
import numpy as np

from .spaces import Space


class MultiDiscrete(Space):
    def __init__(self, array_of_param_array):
        params = np.asarray(array_of_param_array, dtype=np.int64).reshape(-1, 2)
        self.low = params[:, 0].copy()
        self.high = params[:, 1].copy()
        self.num_discrete_space = int(self.low.shape[0])

    @property
    def sizes(self):
        return self.high - self.low + 1

    def sample(self):
        u = np.random.rand(self.num_discrete_space)
        return [int(v) for v in np.floor(self.sizes * u + self.low)]

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == (self.num_discrete_space,) and bool(np.all(x >= self.low) and np.all(x <= self.high))

    @property
    def shape(self):
        return self.num_discrete_space

    def __repr__(self):
        return "MultiDiscrete%d" % self.num_discrete_space

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.low, other.low) and \
            np.array_equal(self.high, other.high)
