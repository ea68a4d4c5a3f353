"""Minimal gym-compatible space types.

The reference builds `gym.spaces.Discrete/Box` objects (environment.py:39-70) only so that a
trainer can read `.n` / `.shape`; `gym` is an optional dependency here.  When gym (or gymnasium)
is importable its classes are used, otherwise these stand-ins expose the same attributes.
"""
import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    from gym import spaces as _gs
    Discrete, Box, Tuple, Space = _gs.Discrete, _gs.Box, _gs.Tuple, _gs.Space
except Exception:  # noqa: BLE001
    class Space(object):
        def contains(self, x):
            raise NotImplementedError

        def sample(self):
            raise NotImplementedError

    class Discrete(Space):
        """{0, ..., n-1}"""

        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return "Discrete(%d)" % self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.full(self.shape, low, dtype=dtype)
            self.high = np.full(self.shape, high, dtype=dtype)
            self.dtype = np.dtype(dtype)

        def sample(self):
            return np.random.uniform(-1.0, 1.0, self.shape).astype(self.dtype)

        def contains(self, x):
            return np.shape(x) == self.shape

        def __repr__(self):
            return "Box%s" % (self.shape,)

    class Tuple(Space):
        def __init__(self, spaces_):
            self.spaces = tuple(spaces_)
