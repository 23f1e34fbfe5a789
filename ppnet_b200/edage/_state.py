"""Process-wide Philox seed and id counters (the replacement for the reference's global RNG streams)."""
import itertools

from ..ops import DEFAULT_SEED

_seed = DEFAULT_SEED
_path_ids = itertools.count()
_gmm_ids = itertools.count()
_map_calls = itertools.count()


def seed(s):
    """Replaces np.random.seed(s) / torch.manual_seed(s): restarts every id counter."""
    global _seed, _path_ids, _gmm_ids, _map_calls
    _seed = int(s) & 0xFFFFFFFFFFFFFFFF
    _path_ids, _gmm_ids, _map_calls = itertools.count(), itertools.count(), itertools.count()


def current_seed():
    return _seed


def next_path_ids(n):
    first = next(_path_ids)
    for _ in range(n - 1):
        next(_path_ids)
    return first


def next_gmm_id():
    return next(_gmm_ids)


def next_map_call():
    return next(_map_calls)
