// discrete_gamma.hpp -- K equiprobable gamma rate categories for the gamma model (host side, O(K) per optimiser step).
//
// Behaviour of the reference's get_gamma (src/gamma.cpp:225-241 -> discrete_gamma :15-52, median branch): the multiplier of
// category i is the median of its slice of Gamma(alpha, alpha), i.e. the chi-square percentage point at (2i+1)/(2K) with
// 2*alpha degrees of freedom divided by 2*alpha, and the K values are rescaled to mean 1; every category has probability 1/K.
// Written from the published algorithms the reference (via PAML) uses: AS 91 (chi-square percentage points, Best & Roberts
// 1975), AS 32 (incomplete gamma integral, Bhattacharjee 1970) and AS 70 (normal percentage points, Odeh & Evans 1974), in
// the same floating-point evaluation order, because the multipliers are INPUTS of the matrix keys (lambda * multiplier is
// truncated to 1e-9, src/matrix_cache.h:49-62): a last-bit difference could move a key.
#pragma once
#include <cmath>
#include <vector>

namespace cafe_b200_host {

inline double point_normal(double prob)
{
    static const double a[5] = {-.322232431088, -1.0, -.342242088547, -.0204231210245, -.453642210148e-4};
    static const double b[5] = {.0993484626060, .588581570495, .531103462366, .103537752850, .0038560700634};
    const double p1 = prob < 0.5 ? prob : 1 - prob;
    if (p1 < 1e-20) return -9999.0;
    const double y = std::sqrt(std::log(1 / (p1 * p1)));
    const double z = y + ((((y * a[4] + a[3]) * y + a[2]) * y + a[1]) * y + a[0]) / ((((y * b[4] + b[3]) * y + b[2]) * y + b[1]) * y + b[0]);
    return prob < 0.5 ? -z : z;
}

// P(alpha, x), regularised lower incomplete gamma: series for small x, continued fraction otherwise
inline double incomplete_gamma(double x, double alpha, double ln_gamma_alpha)
{
    const double accurate = 1e-8, overflow = 1e30;
    const double p = alpha, g = ln_gamma_alpha;
    if (x == 0) return 0.0;
    if (x < 0 || p <= 0) return -1.0;
    const double factor = std::exp(p * std::log(x) - x - g);
    if (!(x > 1 && x >= p)) {
        double gin = 1.0, term = 1.0, rn = p;
        do {
            rn += 1;
            term *= x / rn;
            gin += term;
        } while (term > accurate);
        return gin * (factor / p);
    }
    double a = 1 - p, b = a + x + 1, term = 0.0;
    double pn[6] = {1.0, x, x + 1, x * b, 0.0, 0.0};
    double gin = pn[2] / pn[3];
    for (;;) {
        a += 1;
        b += 2;
        term += 1;
        const double an = a * term;
        pn[4] = b * pn[2] - an * pn[0];
        pn[5] = b * pn[3] - an * pn[1];
        if (pn[5] != 0) {
            const double rn = pn[4] / pn[5];
            const double dif = std::fabs(gin - rn);
            if (dif <= accurate && dif <= accurate * rn) return 1 - factor * gin;
            gin = rn;
        }
        for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
        if (std::fabs(pn[4]) >= overflow)
            for (int i = 0; i < 4; ++i) pn[i] /= overflow;
    }
}

inline double point_chi2(double prob, double v)
{
    const double e = .5e-6, aa = .6931471805;
    const double p = prob;
    if (p < .000002 || p > .999998 || v <= 0) return -1.0;
    const double g = std::lgamma(v / 2);
    const double xx = v / 2, c = xx - 1;
    double ch;
    bool refine = true;
    if (v < -1.24 * std::log(p)) {
        ch = std::pow(p * xx * std::exp(g + xx * aa), 1 / xx);
        if (ch - e < 0) refine = false;
    } else if (v > .32) {
        const double x = point_normal(p);
        const double p1 = 0.222222 / v;
        ch = v * std::pow(x * std::sqrt(p1) + 1 - p1, 3.0);
        if (ch > 2.2 * v + 6) ch = -2 * (std::log(1 - p) - c * std::log(.5 * ch) + g);
    } else {
        ch = 0.4;
        const double a = std::log(1 - p);
        double q;
        do {
            q = ch;
            const double p1 = 1 + ch * (4.67 + ch);
            const double p2 = ch * (6.73 + ch * (6.66 + ch));
            const double t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
            ch -= (1 - std::exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
        } while (std::fabs(q / ch - 1) - .01 > 0);
    }
    if (!refine) return ch;
    double q;
    do {
        q = ch;
        const double p1 = .5 * ch;
        double t = incomplete_gamma(p1, xx, g);
        if (t < 0) return -1.0;
        const double p2 = p - t;
        t = p2 * std::exp(xx * aa + g + p1 - c * std::log(ch));
        const double b = t / ch;
        const double a = 0.5 * t - b * c;
        const double s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
        const double s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
        const double s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
        const double s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
        const double s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
        const double s6 = (120 + c * (346 + 127 * c)) / 5040;
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
    } while (std::fabs(q / ch - 1) > e);
    return ch;
}

// cat_probs[k] = 1/K, multipliers[k] = rescaled category medians
inline void discrete_gamma(int n_cat, double alpha, std::vector<double>& cat_probs, std::vector<double>& multipliers)
{
    cat_probs.assign(n_cat, 1.0 / n_cat);
    multipliers.assign(n_cat, 0.0);
    const double gap05 = 1.0 / (2.0 * n_cat);
    const double factor = alpha / alpha * n_cat;
    double t = 0.0;
    for (int i = 0; i < n_cat; ++i) {
        multipliers[i] = point_chi2((i * 2.0 + 1) * gap05, 2.0 * alpha) / (2.0 * alpha);
        t += multipliers[i];
    }
    for (int i = 0; i < n_cat; ++i) multipliers[i] *= factor / t;
}

}  // namespace cafe_b200_host
