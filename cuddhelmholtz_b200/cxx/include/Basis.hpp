// Basis(n): Lagrange basis on n Gauss-Lobatto nodes (reference include/Basis.hpp:13-63, source/Basis.cpp:109-170).
#ifndef CUDDH_BASIS_HPP
#define CUDDH_BASIS_HPP

#include <algorithm>
#include <memory>

#include "QuadratureRule.hpp"
#include "Tensor.hpp"

namespace cuddh
{
    class Basis
    {
    public:
        Basis(int n_) : n(n_), q(n_, QuadratureRule::GaussLobatto), M(n_, n_), D(n_, n_)
        {
            cuddh_basis_t raw = nullptr;
            cuddh_check(cuddh_b200_basis_create(n, &raw));
            h.reset(raw, [](cuddh_basis_t p) { cuddh_b200_basis_destroy(p); });
            // 1-D mass matrix with the n-point Gauss-Legendre rule and derivative matrix at the nodes (Basis.cpp:118-139)
            QuadratureRule gl(n, QuadratureRule::GaussLegendre);
            dmat P(n, n);
            eval(n, gl.x(), P);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j <= i; ++j) {
                    double m = 0.0;
                    for (int k = 0; k < n; ++k)
                        m += gl.w(k) * P(k, i) * P(k, j);
                    M(i, j) = m;
                    M(j, i) = m;
                }
            deriv(n, q.x(), D);
        }

        int size() const { return n; }

        /// P (m, n) column-major: P(i, j) = j-th basis function at x[i]
        void eval(int m, const double * x, double * P) const { cuddh_check(cuddh_b200_basis_eval(h.get(), m, x, P)); }
        /// D (m, n) column-major: derivative of the j-th basis function at x[i]
        void deriv(int m, const double * x, double * D_) const { cuddh_check(cuddh_b200_basis_deriv(h.get(), m, x, D_)); }

        const_dmat_wrapper mass_matrix() const { return const_dmat_wrapper(M.data(), n, n); }
        const_dmat_wrapper derivative_matrix() const { return const_dmat_wrapper(D.data(), n, n); }
        const QuadratureRule & quadrature() const { return q; }

    private:
        int n;
        QuadratureRule q;
        dmat M, D;
        std::shared_ptr<cuddh_basis_s> h;
    };
} // namespace cuddh

#endif
