// nelder_mead.hpp -- the simplex search CAFE5 runs over its scorers, as a small self-contained minimiser.
//
// Behaviour follows the reference's fminsearch (src/optimizer.cpp:57-73 constants, :163-194 initial simplex, :287-322 main loop,
// :130-161 + :586-589 stopping rule) so that, fed the same objective values, it visits the same points in the same order:
//   * initial simplex: vertex 0 = x0; vertex i = x0 with coordinate i-1 scaled by (1 + delta) (0 -> zero_delta); if the previous
//     vertex scored +inf (and i > 1) the scale is (1 + 100 delta) instead;
//   * vertices kept sorted by score, ties in insertion order (the reference's std::sort runs insertion sort on <= 16 elements);
//   * per iteration: centroid of all but the worst, reflection (rho); expansion (chi) if the reflection beats the best --
//     accepted only if strictly better than the reflection; contraction outside (psi) when the reflection EQUALS the worst
//     score, inside when it is worse; shrink (sigma) towards the best when the contraction fails;
//   * stop when max |x_i+1 - x_i| <= tolx AND max |f_i - f_0| <= tolf (checked before each iteration), or after max_iters.
// Defaults are the reference's: rho 1, chi 2, psi 0.5, sigma 0.5, delta 0.05, zero_delta 0.00025, tol 1e-6, 300 iterations
// (optimizer_parameters::neldermead_iterations, src/optimizer.h:29).
#pragma once
#include <cmath>
#include <functional>
#include <limits>
#include <vector>

namespace cafe_b200_host {

struct NelderMeadOptions {
    double rho = 1.0, chi = 2.0, psi = 0.5, sigma = 0.5;
    double delta = 0.05, zero_delta = 0.00025;
    double tolx = 1e-6, tolf = 1e-6;
    int max_iters = 300;
};

struct NelderMeadResult {
    std::vector<double> x;
    double f = std::numeric_limits<double>::infinity();
    int iterations = 0;
    bool hit_max = false;
};

class NelderMead {
public:
    using Objective = std::function<double(const double*)>;

    NelderMead(Objective f, int n, NelderMeadOptions opt = NelderMeadOptions()) : _f(std::move(f)), _n(n), _opt(opt) {}

    NelderMeadResult minimize(const double* x0)
    {
        const int n = _n;
        _v.assign(n + 1, Vertex{std::vector<double>(n), 0.0});
        for (int i = 0; i <= n; ++i) {
            const bool wide = i > 1 && std::isinf(_v[i - 1].f);
            const double scale = 1 + (wide ? _opt.delta * 100 : _opt.delta);
            for (int j = 0; j < n; ++j)
                _v[i].x[j] = (i - 1 == j) ? (x0[j] ? scale * x0[j] : _opt.zero_delta) : x0[j];
            _v[i].f = _f(_v[i].x.data());
        }
        sort();
        std::vector<double> mean(n), xr(n), xt(n);
        int it = 0;
        for (; it < _opt.max_iters; ++it) {
            if (converged()) break;
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int i = 0; i < n; ++i) s += _v[i].x[j];
                mean[j] = s / n;
            }
            const Vertex& worst = _v[n];
            for (int j = 0; j < n; ++j) xr[j] = mean[j] + _opt.rho * (mean[j] - worst.x[j]);
            const double fr = _f(xr.data());
            if (fr < _v[0].f) {
                for (int j = 0; j < n; ++j) xt[j] = mean[j] + _opt.chi * (xr[j] - mean[j]);
                const double fe = _f(xt.data());
                if (fe < fr) replace_worst(xt, fe);
                else replace_worst(xr, fr);
            } else if (fr >= worst.f) {
                if (fr > worst.f) {
                    for (int j = 0; j < n; ++j) xt[j] = mean[j] + _opt.psi * (mean[j] - worst.x[j]);
                    const double fc = _f(xt.data());
                    if (fc < worst.f) replace_worst(xt, fc);
                    else shrink();
                } else {
                    for (int j = 0; j < n; ++j) xt[j] = mean[j] + _opt.psi * (xr[j] - mean[j]);
                    const double fc = _f(xt.data());
                    if (fc <= fr) replace_worst(xt, fc);
                    else shrink();
                }
            } else {
                replace_worst(xr, fr);
            }
        }
        NelderMeadResult r;
        r.x = _v[0].x;
        r.f = _v[0].f;
        r.iterations = it;
        r.hit_max = it == _opt.max_iters;
        return r;
    }

private:
    struct Vertex {
        std::vector<double> x;
        double f;
    };

    void sort()   // insertion sort: stable, and what std::sort does for a handful of elements
    {
        for (size_t i = 1; i < _v.size(); ++i) {
            Vertex key = std::move(_v[i]);
            size_t j = i;
            while (j > 0 && key.f < _v[j - 1].f) { _v[j] = std::move(_v[j - 1]); --j; }
            _v[j] = std::move(key);
        }
    }

    void replace_worst(const std::vector<double>& x, double f)
    {
        _v[_n].x = x;
        _v[_n].f = f;
        sort();
    }

    void shrink()
    {
        for (int i = 1; i <= _n; ++i) {
            for (int j = 0; j < _n; ++j) _v[i].x[j] = _v[0].x[j] + _opt.sigma * (_v[i].x[j] - _v[0].x[j]);
            _v[i].f = _f(_v[i].x.data());
        }
        sort();
    }

    bool converged() const
    {
        double dx = -std::numeric_limits<double>::max(), df = dx;
        for (int i = 0; i < _n; ++i)
            for (int j = 0; j < _n; ++j) dx = std::fmax(dx, std::fabs(_v[i + 1].x[j] - _v[i].x[j]));
        for (int i = 1; i <= _n; ++i) df = std::fmax(df, std::fabs(_v[i].f - _v[0].f));
        return dx <= _opt.tolx && df <= _opt.tolf;
    }

    Objective _f;
    int _n;
    NelderMeadOptions _opt;
    std::vector<Vertex> _v;
};

}  // namespace cafe_b200_host
