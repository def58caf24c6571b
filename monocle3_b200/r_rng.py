"""R's default RNG stream (Mersenne-Twister + inversion) for the Python host layer.

In the R package the shim keeps `set.seed(2016)` (R/preprocess_cds.R:110) and
lets R draw irlba's start vector `rnorm(n)`; this module plays that part when
the host language is Python.  NumPy's MT19937 bit generator uses the same
tempering as R's MT_genrand, so R's scrambled seed state is injected and the
raw 32-bit stream is converted exactly as R does (unif_rand fixup, two uniforms
per normal with BIG = 2^27, qnorm).
"""
import numpy as np
from scipy.special import ndtri

_I2_32M1 = 2.328306437080797e-10
_BIG = 134217728.0


def _r_seed_state(seed):
    s = np.uint64(seed & 0xFFFFFFFF)
    mult, mask = np.uint64(69069), np.uint64(0xFFFFFFFF)
    for _ in range(50):
        s = (mult * s + np.uint64(1)) & mask
    key = np.empty(625, dtype=np.uint64)
    for j in range(625):
        s = (mult * s + np.uint64(1)) & mask
        key[j] = s
    return key[1:].astype(np.uint32)  # dummy[0] is mti and gets overwritten with 624


def _unif(raw):
    u = raw.astype(np.float64) * 2.3283064365386963e-10
    u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
    return np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)


class RStream:
    """`set.seed(seed)` followed by any number of `rnorm(n)` calls on the same stream (irlba draws
    its start vector first and, on an invariant subspace, replacement vectors later)."""

    def __init__(self, seed):
        self.bg = np.random.MT19937()
        st = self.bg.state
        st["state"]["key"] = _r_seed_state(seed)
        st["state"]["pos"] = 624
        self.bg.state = st

    def rnorm(self, n):
        raw = self.bg.random_raw(2 * int(n)).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        u = _unif(raw)
        a = np.floor(_BIG * u[0::2]) + u[1::2]
        return ndtri(a / _BIG)


def rnorm_after_seed(seed, n):
    """Equivalent of `set.seed(seed); rnorm(n)` in R (default RNG kinds)."""
    return RStream(seed).rnorm(n)
