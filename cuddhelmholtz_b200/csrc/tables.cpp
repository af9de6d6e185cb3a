// 1-D numerics: Gauss-Legendre / Gauss-Lobatto rules and the Lagrange basis on GLL nodes.
//
// Mirrors the semantics of the reference's QuadratureRule (source/QuadratureRule.cpp:64-202) and
// Basis (source/Basis.cpp:3-170): same literal node tables for small n, same weight formulas through
// Jacobi-polynomial recurrences and lgamma/exp, same barycentric normalisation and node short-circuit,
// so the tables P(nq,nb), D(nq,nb) fed to the CUDA kernels are bit-identical to the reference's for
// every n the literal tables cover (GL n<=10, GLL n<=9). Larger rules use Golub-Welsch: the reference
// calls LAPACK dsteqr_ (un-vendored); here the eigenvalues come from a self-contained implicit-QL
// iteration and are polished by the same three Newton steps (agreement <= 1 ulp, tested).
//
// This translation unit must be compiled WITHOUT floating-point contraction (-ffp-contract=off).
#include "common.hpp"
#include <algorithm>
#include <cmath>
#include <limits>

namespace cb200
{
    namespace
    {
        // literal nodes, digits as published in the reference tables (:72-83, :142-151)
        const double GL_NODES[11][10] = {
            {0},
            {0.0},
            {-0.577350269189625764509149, 0.577350269189625764509149},
            {-0.774596669241483377035853, 0.0, 0.774596669241483377035853},
            {-0.861136311594052575223946, -0.339981043584856264802666, 0.339981043584856264802666, 0.861136311594052575223946},
            {-0.906179845938663992797627, -0.538469310105683091036314, 0.0, 0.538469310105683091036314, 0.906179845938663992797627},
            {-0.932469514203152027812302, -0.661209386466264513661400, -0.238619186083196908630502, 0.238619186083196908630502, 0.661209386466264513661400, 0.932469514203152027812302},
            {-0.949107912342758524526190, -0.741531185599394439863865, -0.405845151377397166906606, 0.0, 0.405845151377397166906606, 0.741531185599394439863865, 0.949107912342758524526190},
            {-0.960289856497536231683561, -0.796666477413626739591554, -0.525532409916328985817739, -0.183434642495649804939476, 0.183434642495649804939476, 0.525532409916328985817739, 0.796666477413626739591554, 0.960289856497536231683561},
            {-0.968160239507626089835576, -0.836031107326635794299430, -0.613371432700590397308702, -0.324253423403808929038538, 0.0, 0.324253423403808929038538, 0.613371432700590397308702, 0.836031107326635794299430, 0.968160239507626089835576},
            {-0.973906528517171720077964, -0.865063366688984510732097, -0.679409568299024406234327, -0.433395394129247190799266, -0.148874338981631210884826, 0.148874338981631210884826, 0.433395394129247190799266, 0.679409568299024406234327, 0.865063366688984510732097, 0.973906528517171720077964}};

        const double GLL_NODES[10][9] = {
            {0},
            {0},
            {-1, 1},
            {-1, 0, 1},
            {-1, -0.447213595499958, 0.447213595499958, 1},
            {-1, -0.654653670707977, 0, 0.654653670707977, 1},
            {-1, -0.765055323929465, -0.285231516480645, 0.285231516480645, 0.765055323929465, 1},
            {-1, -0.830223896278567, -0.468848793470714, 0.0, 0.468848793470714, 0.830223896278567, 1},
            {-1, -0.871740148509607, -0.591700181433142, -0.209299217902479, 0.2092992179024789, 0.591700181433142, 0.871740148509607, 1},
            {-1, -0.899757995411460, -0.677186279510738, -0.363117463826178, 0, 0.363117463826178, 0.677186279510738, 0.899757995411460, 1}};

        // Jacobi polynomial P_n^{(a,b)}(x) by the three-term recurrence (reference :21-46)
        double jacobi(unsigned n, double a, double b, double x)
        {
            double prev = 1;
            if (n == 0)
                return prev;
            double cur = (a + 1) + 0.5 * (a + b + 2) * (x - 1);
            for (unsigned m = 2; m <= n; ++m) {
                double next = (2 * m + a + b - 1) * ((2 * m + a + b) * (2 * m + a + b - 2) * x + a * a - b * b) * cur -
                              2 * (m + a - 1) * (m + b - 1) * (2 * m + a + b) * prev;
                next /= 2 * m * (m + a + b) * (2 * m + a + b - 2);
                prev = cur;
                cur = next;
            }
            return cur;
        }

        // k-th derivative of P_n^{(a,b)} (reference :48-57)
        double jacobi_deriv(unsigned k, unsigned n, double a, double b, double x)
        {
            if (k > n)
                return 0.0;
            const double s = std::lgamma(n + a + b + 1 + k) - std::lgamma(n + a + b + 1) - k * std::log(2);
            return std::exp(s) * jacobi(n - k, a + k, b + k, x);
        }

        inline double sq(double v) { return v * v; }

        // eigenvalues (ascending) of the symmetric tridiagonal matrix (d, e): implicit QL with Wilkinson shifts
        void tridiag_eigenvalues(int n, double * d, double * e /* n-1, destroyed; needs n slots */)
        {
            if (n <= 1)
                return;
            e[n - 1] = 0.0;
            for (int l = 0; l < n; ++l) {
                int iter = 0;
                int m;
                do {
                    for (m = l; m < n - 1; ++m) {
                        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                        if (std::fabs(e[m]) <= std::numeric_limits<double>::epsilon() * dd)
                            break;
                    }
                    if (m != l) {
                        CB_REQUIRE(iter++ < 200, "QuadratureRule: tridiagonal eigenvalue iteration did not converge");
                        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                        double r = std::hypot(g, 1.0);
                        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                        double s = 1.0, c = 1.0, p = 0.0;
                        int i;
                        for (i = m - 1; i >= l; --i) {
                            double f = s * e[i];
                            const double b = c * e[i];
                            e[i + 1] = (r = std::hypot(f, g));
                            if (r == 0.0) {
                                d[i + 1] -= p;
                                e[m] = 0.0;
                                break;
                            }
                            s = f / r;
                            c = g / r;
                            g = d[i + 1] - p;
                            r = (d[i] - g) * s + 2.0 * c * b;
                            d[i + 1] = g + (p = s * r);
                            g = c * r - b;
                        }
                        if (r == 0.0 && i >= l)
                            continue;
                        d[l] -= p;
                        e[l] = g;
                        e[m] = 0.0;
                    }
                } while (m != l);
            }
            std::sort(d, d + n);
        }

        void gauss_legendre(int n, double * x, double * w)
        {
            CB_REQUIRE(n >= 1, "QuadratureRule error: Guass-Legendre rules require n >= 1");
            if (n <= 10) {
                for (int i = 0; i < n; ++i)
                    x[i] = GL_NODES[n][i];
            }
            else { // Golub-Welsch + 3 Newton steps on the lower half, mirrored (reference :93-127)
                std::fill_n(x, n, 0.0);
                std::vector<double> E(n);
                for (int i = 0; i < n - 1; ++i) {
                    const double k = i + 1;
                    E[i] = k * std::sqrt(1.0 / (4.0 * k * k - 1.0));
                }
                tridiag_eigenvalues(n, x, E.data());
                for (int i = 0; i < n / 2; ++i) {
                    for (int j = 0; j < 3; ++j) {
                        const double P = jacobi(n, 0, 0, x[i]);
                        const double dP = jacobi_deriv(1, n, 0, 0, x[i]);
                        x[i] -= P / dP;
                    }
                    x[n - 1 - i] = -x[i];
                }
                if (n & 1)
                    x[n / 2] = 0.0;
            }
            for (int i = 0; i < n; ++i)
                w[i] = 2.0 / (1.0 - sq(x[i])) / sq(jacobi_deriv(1, n, 0, 0, x[i]));
        }

        void gauss_lobatto(int n, double * x, double * w)
        {
            CB_REQUIRE(n >= 2, "QuadratureRule error: Gauss-Lobatto rules require n >= 2");
            if (n <= 9) {
                for (int i = 0; i < n; ++i)
                    x[i] = GLL_NODES[n][i];
            }
            else { // reference :158-197
                std::fill_n(x + 1, n - 2, 0.0);
                std::vector<double> E(n);
                for (int i = 0; i < n - 3; ++i) {
                    const double ii = i + 1;
                    E[i] = std::sqrt(ii * (ii + 2.0) / ((2.0 * ii + 3.0) * (2.0 * ii + 1.0)));
                }
                tridiag_eigenvalues(n - 2, x + 1, E.data());
                x[0] = -1.0;
                x[n - 1] = 1.0;
                for (int i = 1; i < n / 2; ++i) {
                    for (int j = 0; j < 3; ++j) {
                        const double P = jacobi(n - 2, 1, 1, x[i]);
                        const double dP = jacobi_deriv(1, n - 2, 1, 1, x[i]);
                        x[i] -= P / dP;
                    }
                    x[n - 1 - i] = -x[i];
                }
                if (n & 1)
                    x[n / 2] = 0.0;
            }
            for (int i = 0; i < n; ++i)
                w[i] = 2.0 / (n * (n - 1) * sq(jacobi(n - 1, 0, 0, x[i])));
        }
    } // namespace

    void quadrature_rule(int n, int type, double * x, double * w)
    {
        if (type == GAUSS_LEGENDRE)
            gauss_legendre(n, x, w);
        else
            gauss_lobatto(n, x, w);
    }

    // ---- Lagrange basis on GLL nodes, barycentric form (reference source/Basis.cpp) ----
    Basis::Basis(int n_) : n(n_), x(n_), w(n_), wb(n_)
    {
        CB_REQUIRE(n >= 2, "Basis requires n >= 2");
        gauss_lobatto(n, x.data(), w.data());
        for (int i = 0; i < n; ++i) { // :3-24
            double t = 1.0;
            for (int j = 0; j < n; ++j)
                if (i != j)
                    t *= x[i] - x[j];
            wb[i] = 1.0 / t;
        }
        const auto mm = std::minmax_element(wb.begin(), wb.end());
        const double diff = *mm.second - *mm.first;
        for (int i = 0; i < n; ++i)
            wb[i] /= diff;
    }

    namespace
    {
        // value at x0 of the interpolant with nodal values y = e_sel (reference :32-52)
        double lagrange_value(const Basis & b, double x0, int sel)
        {
            constexpr double eps = std::numeric_limits<double>::epsilon();
            double A = 0.0, B = 0.0;
            for (int i = 0; i < b.n; ++i) {
                const double xdiff = x0 - b.x[i];
                if (x0 == b.x[i] || std::abs(xdiff) <= eps)
                    return (i == sel) ? 1.0 : 0.0;
                const double C = b.wb[i] / xdiff;
                A += C * ((i == sel) ? 1.0 : 0.0);
                B += C;
            }
            return A / B;
        }

        // derivative at x0 (reference :60-105)
        double lagrange_slope(const Basis & b, double x0, int sel)
        {
            constexpr double eps = std::numeric_limits<double>::epsilon();
            const int n = b.n;
            double A = 0.0, B = 0.0;
            const double p = lagrange_value(b, x0, sel);
            bool atnode = false;
            int inode = -1;
            for (int j = 0; j < n; ++j)
                if (x0 == b.x[j] || std::abs(x0 - b.x[j]) <= eps) {
                    atnode = true;
                    B = -b.wb[j];
                    inode = j;
                }
            if (atnode) {
                for (int j = 0; j < n; ++j) {
                    if (j == inode)
                        continue;
                    const double yj = (j == sel) ? 1.0 : 0.0;
                    A += b.wb[j] * (p - yj) / (x0 - b.x[j]);
                }
            }
            else {
                for (int j = 0; j < n; ++j) {
                    const double yj = (j == sel) ? 1.0 : 0.0;
                    const double t = b.wb[j] / (x0 - b.x[j]);
                    A += t * (p - yj) / (x0 - b.x[j]);
                    B += t;
                }
            }
            return A / B;
        }
    } // namespace

    void Basis::eval(int m, const double * xq, double * P) const
    {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j)
                P[j + (size_t)m * i] = lagrange_value(*this, xq[j], i);
    }

    void Basis::deriv(int m, const double * xq, double * D) const
    {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j)
                D[j + (size_t)m * i] = lagrange_slope(*this, xq[j], i);
    }
} // namespace cb200
