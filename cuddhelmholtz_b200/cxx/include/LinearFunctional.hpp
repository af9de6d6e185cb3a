// LinearFunctional: F[i] += c * (f, phi[i]) for a user function f(x) given as a __device__ callable
// (reference include/LinearFunctional.hpp). The user's lambda runs in one flat kernel that evaluates
// w_i w_j detJ f(x_ij) straight from the element corners (no cached metric arrays); the contraction with the basis and
// the ordered, atomic-free assembly happen in the library (csrc/aux.cu).
#ifndef CUDDH_LINEAR_FUNCTIONAL_HPP
#define CUDDH_LINEAR_FUNCTIONAL_HPP

#include "H1Space.hpp"
#include "forall.hpp"
#include "linalg.hpp"

namespace cuddh
{
    class LinearFunctional
    {
    public:
        /// collocated rule on the basis' own Gauss-Lobatto nodes
        LinearFunctional(const H1Space & fem_)
            : fem(fem_), ndof(fem_.size()), n_elem(fem_.mesh().n_elem()), n_basis(fem_.basis().size()), n_quad(n_basis), fast(true),
              _x(n_quad), _w(n_quad)
        {
            load_rule(fem.basis().quadrature());
        }

        LinearFunctional(const H1Space & fem_, const QuadratureRule & quad)
            : fem(fem_), ndof(fem_.size()), n_elem(fem_.mesh().n_elem()), n_basis(fem_.basis().size()), n_quad(quad.size()), fast(false),
              _x(n_quad), _w(n_quad), _P(n_quad * n_basis)
        {
            load_rule(quad);
            fem.basis().eval(n_quad, quad.x(), _P.host_write());
        }

        /// F[i] <- F[i] + c * (f, phi[i]);  f: (const double x[2]) -> double, __device__
        template <typename Func>
        void action(double c, Func && f, double * F) const
        {
            const int nq = n_quad;
            const long long n_pts = (long long)nq * nq * n_elem;
            if (n_pts >= 2147483647LL)
                cuddh_error("LinearFunctional::action: too many quadrature points for one launch");
            _g.resize((int)n_pts);
            double * g = _g.device_write();
            const double * xq = _x.device_read();
            const double * wq = _w.device_read();
            const double * corners = cuddh_b200_h1space_device_corners(fem.handle());
            forall((int)n_pts, [=] __device__(int t) -> void {
                const int el = t / (nq * nq);
                const int rem = t - el * nq * nq;
                const int i = rem % nq, j = rem / nq;
                const double * cr = corners + 8 * (size_t)el;
                const double xi0 = xq[i], xi1 = xq[j];
                const double b[4] = {0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1),
                                     0.25 * (1.0 + xi0) * (1.0 + xi1), 0.25 * (1.0 - xi0) * (1.0 + xi1)};
                double X[2] = {0.0, 0.0};
                for (int k = 0; k < 4; ++k) {
                    X[0] += cr[2 * k] * b[k];
                    X[1] += cr[2 * k + 1] * b[k];
                }
                const double J0 = 0.25 * ((1.0 - xi1) * (cr[2] - cr[0]) + (1.0 + xi1) * (cr[4] - cr[6]));
                const double J1 = 0.25 * ((1.0 - xi1) * (cr[3] - cr[1]) + (1.0 + xi1) * (cr[5] - cr[7]));
                const double J2 = 0.25 * ((1.0 - xi0) * (cr[6] - cr[0]) + (1.0 + xi0) * (cr[4] - cr[2]));
                const double J3 = 0.25 * ((1.0 - xi0) * (cr[7] - cr[1]) + (1.0 + xi0) * (cr[5] - cr[3]));
                g[t] = wq[i] * wq[j] * (J0 * J3 - J1 * J2) * f(X);
            });
            cuddh_check(cuddh_b200_linear_functional_assemble(fem.handle(), nq, fast ? nullptr : _P.host_read(), g, c, F, nullptr));
        }

        /// F[i] <- (f, phi[i])
        template <typename Func>
        void action(Func && f, double * F) const
        {
            zeros(ndof, F);
            action(1.0, f, F);
        }

    private:
        void load_rule(const QuadratureRule & q)
        {
            double * hx = _x.host_write();
            double * hw = _w.host_write();
            for (int i = 0; i < n_quad; ++i) {
                hx[i] = q.x(i);
                hw[i] = q.w(i);
            }
        }

        const H1Space & fem;
        const int ndof, n_elem, n_basis, n_quad;
        const bool fast;
        host_device_dvec _x, _w, _P;
        mutable host_device_dvec _g;
    };
} // namespace cuddh

#endif
