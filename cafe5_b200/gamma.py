"""Discrete gamma rate categories for the gamma model (host side; O(K) scalar work per optimiser step).

Behaviour of the reference's get_gamma (src/gamma.cpp:225-241 -> discrete_gamma :15-52, median branch):
K equiprobable categories whose multipliers are the category medians of Gamma(alpha, alpha), rescaled to
mean 1.  Uses the published algorithms AS 91 (chi-square percentage points; Best & Roberts 1975), AS 32
(incomplete gamma integral; Bhattacharjee 1970) and AS 70 (normal percentage points; Odeh & Evans 1974),
the ones PAML (and through it the reference) uses, so the multipliers agree to the last bit with a libm
that matches (they are inputs of the kernel, so any difference would shift every matrix key).
"""
import ctypes
import ctypes.util
import math

# CPython's math.lgamma is its own Lanczos implementation; the reference calls the C library's lgamma.
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.lgamma.restype = ctypes.c_double
_libm.lgamma.argtypes = [ctypes.c_double]


def libm_lgamma(x):
    return _libm.lgamma(x)


def _point_normal(prob):
    a = (-.322232431088, -1.0, -.342242088547, -.0204231210245, -.453642210148e-4)
    b = (.0993484626060, .588581570495, .531103462366, .103537752850, .0038560700634)
    p1 = prob if prob < 0.5 else 1 - prob
    if p1 < 1e-20:
        return -9999.0
    y = math.sqrt(math.log(1 / (p1 * p1)))
    z = y + ((((y * a[4] + a[3]) * y + a[2]) * y + a[1]) * y + a[0]) / ((((y * b[4] + b[3]) * y + b[2]) * y + b[1]) * y + b[0])
    return -z if prob < 0.5 else z


def _incomplete_gamma(x, alpha, ln_gamma_alpha):
    accurate, overflow = 1e-8, 1e30
    p, g = alpha, ln_gamma_alpha
    if x == 0:
        return 0.0
    if x < 0 or p <= 0:
        return -1.0
    factor = math.exp(p * math.log(x) - x - g)
    if not (x > 1 and x >= p):
        gin = term = 1.0
        rn = p
        while True:
            rn += 1
            term *= x / rn
            gin += term
            if not term > accurate:
                break
        return gin * (factor / p)
    a = 1 - p
    b = a + x + 1
    term = 0.0
    pn = [1.0, x, x + 1, x * b, 0.0, 0.0]
    gin = pn[2] / pn[3]
    while True:
        a += 1
        b += 2
        term += 1
        an = a * term
        pn[4] = b * pn[2] - an * pn[0]
        pn[5] = b * pn[3] - an * pn[1]
        if pn[5] != 0:
            rn = pn[4] / pn[5]
            dif = abs(gin - rn)
            if dif <= accurate and dif <= accurate * rn:
                return 1 - factor * gin
            gin = rn
        pn[0:4] = pn[2:6]
        if abs(pn[4]) >= overflow:
            pn[0:4] = [v / overflow for v in pn[0:4]]


def _point_chi2(prob, v):
    e, aa = .5e-6, .6931471805
    p = prob
    if p < .000002 or p > .999998 or v <= 0:
        return -1.0
    g = libm_lgamma(v / 2)
    xx = v / 2
    c = xx - 1
    if v < -1.24 * math.log(p):
        ch = math.pow(p * xx * math.exp(g + xx * aa), 1 / xx)
        if ch - e < 0:
            return ch
    elif v > .32:
        x = _point_normal(p)
        p1 = 0.222222 / v
        ch = v * math.pow(x * math.sqrt(p1) + 1 - p1, 3.0)
        if ch > 2.2 * v + 6:
            ch = -2 * (math.log(1 - p) - c * math.log(.5 * ch) + g)
    else:
        ch = 0.4
        a = math.log(1 - p)
        while True:
            q = ch
            p1 = 1 + ch * (4.67 + ch)
            p2 = ch * (6.73 + ch * (6.66 + ch))
            t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2
            ch -= (1 - math.exp(a + g + .5 * ch + c * aa) * p2 / p1) / t
            if abs(q / ch - 1) - .01 <= 0:
                break
    while True:
        q = ch
        p1 = .5 * ch
        t = _incomplete_gamma(p1, xx, g)
        if t < 0:
            return -1.0
        p2 = p - t
        t = p2 * math.exp(xx * aa + g + p1 - c * math.log(ch))
        b = t / ch
        a = 0.5 * t - b * c
        s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420
        s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520
        s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520
        s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040
        s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520
        s6 = (120 + c * (346 + 127 * c)) / 5040
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))))
        if not abs(q / ch - 1) > e:
            return ch


def get_gamma(n_cat, alpha):
    """Return (cat_probs, multipliers), each a list of n_cat doubles."""
    gap05 = 1.0 / (2.0 * n_cat)
    factor = alpha / alpha * n_cat
    rates = [_point_chi2((i * 2.0 + 1) * gap05, 2.0 * alpha) / (2.0 * alpha) for i in range(n_cat)]
    t = 0.0
    for r in rates:
        t += r
    rates = [r * (factor / t) for r in rates]   # reference: rK[i] *= factor/t
    return [1.0 / n_cat] * n_cat, rates
