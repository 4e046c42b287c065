"""R-compatible random streams on the host (numpy): set.seed(), runif(), rnorm().

Why it exists: the reference draws from R's global RNG (default Mersenne-Twister + Inversion) in a
fixed, closed-form order: default inits (R/main.R:299-314), then per chain, per
iteration, per rung, per free parameter 2 uniforms (R::rnorm, src/Particle.cpp:49) + 1 (R::runif,
src/Particle.h:139), then one per coupling pair (src/main.cpp:311).  Generating that stream here and
handing it to the GPU as replay uniforms makes a GPU run reproduce, chain for chain, what the
reference prints after the same set.seed() -- the strongest drop-in evidence available.

Host-side helper only (numpy); nothing here runs on the sampling hot path.
"""
from __future__ import annotations

import numpy as np

_I2_32M1 = 2.328306437080797e-10


class RStream:
    """The state of R's default RNG after set.seed(seed)."""

    def __init__(self, seed: int):
        s = np.uint32(seed & 0xFFFFFFFF)
        key = np.empty(625, dtype=np.uint32)
        with np.errstate(over="ignore"):
            for _ in range(50):                      # initial scrambling
                s = np.uint32(69069) * s + np.uint32(1)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                key[j] = s
        # word 0 is the position (forced to 624 = "regenerate on first draw"), words 1..624 the state
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = key[1:].copy()
        st["state"]["pos"] = 624
        self._bg.state = st

    def runif(self, n: int) -> np.ndarray:
        """unif_rand(): 32-bit word * 2^-32, never exactly 0 or 1 (R's fixup)."""
        w = self._bg.random_raw(int(n)).astype(np.float64)
        u = w * 2.3283064365386963e-10
        u[u <= 0.0] = 0.5 * _I2_32M1
        u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * _I2_32M1
        return u

    def rnorm(self, n: int, mean: float = 0.0, sd: float = 1.0) -> np.ndarray:
        """norm_rand() with normal.kind = "Inversion": two uniforms per draw, then qnorm."""
        u = self.runif(2 * int(n)).reshape(-1, 2)
        big = 134217728.0
        p = (np.floor(big * u[:, 0]) + u[:, 1]) / big
        return mean + sd * qnorm(p)


def qnorm(p) -> np.ndarray:
    """Wichura's AS241 (PPND16), the algorithm behind R's qnorm."""
    p = np.asarray(p, dtype=np.float64)
    q = p - 0.5
    out = np.empty_like(p)
    cen = np.abs(q) <= 0.425
    r = 0.180625 - q[cen] * q[cen]
    num = (((((((2509.0809287301226727 * r + 33430.575583588128105) * r + 67265.770927008700853) * r
               + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
            + 133.14166789178437745) * r + 3.387132872796366608)
    den = (((((((5226.495278852545925 * r + 28729.085735721942674) * r + 39307.89580009271061) * r
               + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
            + 42.313330701600911252) * r + 1.0)
    out[cen] = q[cen] * num / den
    tail = ~cen
    pt, qt = p[tail], q[tail]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.sqrt(-np.log(np.where(qt < 0, pt, 1.0 - pt)))
    val = np.empty_like(r)
    mid = r <= 5.0
    rm = r[mid] - 1.6
    val[mid] = ((((((((7.7454501427834140764e-4 * rm + 0.0227238449892691845833) * rm + 0.24178072517745061177) * rm
                     + 1.27045825245236838258) * rm + 3.64784832476320460504) * rm + 5.7694972214606914055) * rm
                  + 4.6303378461565452959) * rm + 1.42343711074968357734)
                / (((((((1.05075007164441684324e-9 * rm + 5.475938084995344946e-4) * rm + 0.0151986665636164571966) * rm
                       + 0.14810397642748007459) * rm + 0.68976733498510000455) * rm + 1.6763848301838038494) * rm
                    + 2.05319162663775882187) * rm + 1.0))
    rf = r[~mid] - 5.0
    val[~mid] = ((((((((2.01033439929228813265e-7 * rf + 2.71155556874348757815e-5) * rf + 0.0012426609473880784386) * rf
                      + 0.026532189526576123093) * rf + 0.29656057182850489123) * rf + 1.7848265399172913358) * rf
                   + 5.4637849111641143699) * rf + 6.6579046435011037772)
                 / (((((((2.04426310338993978564e-15 * rf + 1.4215117583164458887e-7) * rf + 1.8463183175100546818e-5) * rf
                        + 7.868691311456132591e-4) * rf + 0.0148753612908506148525) * rf + 0.13692988092273580531) * rf
                     + 0.59983220655588793769) * rf + 1.0))
    out[tail] = np.where(qt < 0, -val, val)
    return out
