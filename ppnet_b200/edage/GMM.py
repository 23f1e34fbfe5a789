"""EDaGe-PP/GMM.py:7-16 -- random 2-D Gaussian mixture; `.Distribution.sample([N])`."""
import torch

from .. import ops
from . import _state


class _Mixture:
    """Stands in for torch.distributions.MixtureSameFamily: only `.sample(shape)` is used by the reference."""

    def __init__(self, seed, mean, std, weights):
        self._seed, self.mean, self.std, self.weights = seed, mean, std, weights
        self._next = 0

    def sample(self, sample_shape=()):
        shape = list(sample_shape) if not isinstance(sample_shape, int) else [sample_shape]
        n = 1
        for s in shape:
            n *= int(s)
        out = ops.gmm_sample(self._seed, self._next, n, self.mean, self.std, self.weights)
        self._next += n                                   # successive calls continue the counter stream
        return out.reshape(shape + [self.mean.shape[1]])


class GMM:
    def __init__(self, order=10, dim=2, mean_range=70, std_range=5):
        self.Order = order
        self.Dim = dim
        # one Philox key per GMM object: (global seed, object id)
        seed = (_state.current_seed() + 0x9E3779B97F4A7C15 * (1 + _state.next_gmm_id())) & 0xFFFFFFFFFFFFFFFF
        mean, std, weights = ops.gmm_params(seed, order, dim, float(mean_range), float(std_range),
                                            device=torch.device("cuda", torch.cuda.current_device()))
        self.Distribution = _Mixture(seed, mean, std, weights)
