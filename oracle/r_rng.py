"""Restatement of base R's default RNG: set.seed() + runif()/rnorm().

TEST INFRASTRUCTURE ONLY (oracle/).  The reference fixes irlba's start vector
with `set.seed(2016)` (R/preprocess_cds.R:110, R/projection.R:100) and irlba's R
wrapper then draws `rnorm(ncol(A))`.  Base R is not in /root/reference (it is
the language runtime), so this restates its published algorithm:
  * RNGkind "Mersenne-Twister" with R's seed scrambling (RNG.c: Randomize /
    RNG_Init: seed = 69069*seed+1 fifty times, then once per state word),
  * MT19937 tempering and `fixup` into (0,1),
  * normal.kind "Inversion": u = unif_rand(); u = (int)(BIG*u) + unif_rand();
    qnorm5(u/BIG) with BIG = 2^27 (snorm.c).
Pinned by known R outputs:
  set.seed(42);   runif(3) -> 0.9148060 0.9370754 0.2861395
  set.seed(42);   rnorm(3) -> 1.37095845 -0.56469817 0.36312841
  set.seed(2016); rnorm(5) -> -0.91474184 1.00124785 -0.05642291 0.29664516 -2.79147086
qnorm uses scipy.special.ndtri (Cephes) instead of R's AS241; the two agree to
~1e-16 relative, far inside every tolerance on this path.
"""
import numpy as np
from scipy.special import ndtri

_N = 624
_M = 397
_I2_32M1 = 2.328306437080797e-10  # = 1/(2^32 - 1)
_BIG = 134217728.0  # 2^27


class RRandom:
    def __init__(self, seed):
        self.set_seed(seed)

    def set_seed(self, seed):
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
            st = np.empty(_N + 1, dtype=np.uint32)
            for j in range(_N + 1):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
                st[j] = s
        # FixupSeeds: dummy[0] = mti = N  -> forces regeneration on first draw
        self.mt = st[1:].astype(np.uint64)  # keep in uint64 to do masked 32-bit math
        self.mti = _N

    def _genrand_block(self):
        mt = [int(v) for v in self.mt]
        up, lo = 0x80000000, 0x7FFFFFFF
        for kk in range(_N - _M):
            y = (mt[kk] & up) | (mt[kk + 1] & lo)
            mt[kk] = mt[kk + _M] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        for kk in range(_N - _M, _N - 1):
            y = (mt[kk] & up) | (mt[kk + 1] & lo)
            mt[kk] = mt[kk + (_M - _N)] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        y = (mt[_N - 1] & up) | (mt[0] & lo)
        mt[_N - 1] = mt[_M - 1] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.mt = np.array(mt, dtype=np.uint64)
        self.mti = 0

    def unif_rand(self):
        if self.mti >= _N:
            self._genrand_block()
        y = int(self.mt[self.mti])
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        y &= 0xFFFFFFFF
        v = y * 2.3283064365386963e-10
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def runif(self, n):
        return np.array([self.unif_rand() for _ in range(n)])

    def rnorm(self, n):
        u = np.empty(n)
        for k in range(n):
            a = self.unif_rand()
            a = float(int(_BIG * a)) + self.unif_rand()
            u[k] = a / _BIG
        return ndtri(u)


def rnorm_after_seed(seed, n):
    """`set.seed(seed); rnorm(n)`."""
    return RRandom(seed).rnorm(n)
