import numpy as np


def np_random(seed=None):
    """gym 0.13: (numpy RandomState, seed)."""
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (2 ** 31))
    rng = np.random.RandomState()
    rng.seed(int(seed) % (2 ** 32))
    return rng, seed
