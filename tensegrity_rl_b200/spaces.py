"""`Box` space: gymnasium's / gym's when importable (so SB3's space-equality check in SAC.load passes),
otherwise a minimal stand-in with the same fields (this image ships neither package)."""
import numpy as np

try:  # pragma: no cover - depends on the host image
    from gymnasium.spaces import Box  # type: ignore
except Exception:  # noqa: BLE001
    try:
        from gym.spaces import Box  # type: ignore
    except Exception:  # noqa: BLE001

        class Box:  # minimal stand-in
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.shape(low)
                self.shape = tuple(shape)
                self._shape = self.shape
                self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                return self._rng.uniform(lo, hi).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __eq__(self, other):
                return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                        and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
