"""TEST INFRASTRUCTURE ONLY -- how normal is the Walsh-Hadamard generator (STAG_NOISE_NORMAL_HADAMARD)?

z = sum of 128 iid masked FP8 values / sd.  The exact law of that sum is computed here by FFT convolution of the
single-term probability mass function on its lattice (multiples of 2^-8) and compared with the normal CDF; the
script also reruns the search that picked the mask (AND 0xCD, OR 0x12 on e4m3 bytes): among all masks that keep
the sign bit random, the one whose 128-fold sum has the smallest Kolmogorov distance to N(0,1).

    python oracle/wh_quality.py            # the chosen mask: KS distance 5.6e-6, kurtosis 3.0008
    python oracle/wh_quality.py --search   # the full search (about a minute)
"""
import sys

import numpy as np
from scipy.stats import norm


def e4m3(b):
    s = -1.0 if b & 0x80 else 1.0
    e, m = (b >> 3) & 0xF, b & 7
    if e == 0xF and m == 7:
        return np.nan
    if e == 0:
        return s * (m / 8) * 2.0 ** (-6)
    return s * (1 + m / 8) * 2.0 ** (e - 7)


TAB = np.array([e4m3(b) for b in range(256)])


def values(and_mask, or_mask):
    return np.array([TAB[(r & and_mask) | or_mask] for r in range(256)])


def ks_distance(vals, n=128):
    """max |F_n - Phi| of the standardised n-fold sum of iid draws from `vals` (exact up to FFT rounding)."""
    vals = np.asarray(vals, dtype=np.float64)
    step = 2.0 ** np.floor(np.log2(np.abs(vals[vals != 0]).min()))
    while not np.allclose(vals / step, np.round(vals / step)):
        step /= 2
    iv = np.round(vals / step).astype(np.int64)
    size = 1
    while size < 2 * n * int(np.abs(iv).max()) + 2:
        size *= 2
    if size > 1 << 25:
        return None
    pmf = np.zeros(size)
    for v in iv:
        pmf[v % size] += 1.0 / len(iv)
    p = np.fft.irfft(np.fft.rfft(pmf) ** n, size)
    p = np.roll(np.maximum(p, 0), size // 2)
    x = (np.arange(size) - size // 2) * step
    sd = np.sqrt(n * (vals ** 2).mean())
    cdf, phi = np.cumsum(p), norm.cdf(x / sd)
    return max(np.abs(cdf - phi).max(), np.abs(cdf - p - phi).max())


def main():
    v = values(0xCD, 0x12)
    m2, m4 = (v ** 2).mean(), (v ** 4).mean()
    print("mask AND 0xCD OR 0x12: E v^2 = %.10g  kurtosis %.5f  KS distance of the 128-fold sum %.3g"
          % (m2, m4 / m2 ** 2, ks_distance(v)))
    print("for comparison: uniform int8 bytes %.3g" % ks_distance(np.arange(-128, 128) + 0.5))
    if "--search" not in sys.argv:
        return
    cands = []
    for and_mask in range(0x80, 256):
        fixed = [i for i in range(7) if not (and_mask >> i & 1)]
        for bits in range(1 << len(fixed)):
            or_mask = sum(1 << i for j, i in enumerate(fixed) if bits >> j & 1)
            vals = values(and_mask, or_mask)
            if np.isnan(vals).any() or not (vals ** 2).mean():
                continue
            cands.append((abs((vals ** 4).mean() / (vals ** 2).mean() ** 2 - 3), and_mask, or_mask, vals))
    cands.sort(key=lambda t: t[0])
    out = []
    for _, a, o, vals in cands[:60]:
        d = ks_distance(vals)
        if d is not None:
            out.append((d, a, o))
    for d, a, o in sorted(out)[:10]:
        print("KS %.3g  AND 0x%02X OR 0x%02X" % (d, a, o))


if __name__ == "__main__":
    main()
