// gmres(): restarted GMRES(m) with the reference's signatures and solver_out (include/gmres.hpp:14-36). The Krylov
// loop lives in the library (csrc/linalg.cu) and calls back through Operator::action (virtual, device pointers), so
// user-defined operators work unchanged.
#ifndef CUDDH_GMRES_HPP
#define CUDDH_GMRES_HPP

#include <chrono>
#include <iomanip>
#include <iostream>
#include <vector>

#include "Operator.hpp"
#include "Tensor.hpp"
#include "linalg.hpp"

namespace cuddh
{
    struct solver_out
    {
        bool success;
        int num_iter;
        int num_matvec;
        std::vector<double> res_norm;
        std::vector<double> time;
    };

    namespace detail
    {
        // Operator::action has no stream argument (include/Operator.hpp:12-16): user operators enqueue on the legacy default
        // stream, which is also the stream handed to the library below (nullptr), so the ordering is the reference's.
        inline int apply_double(void * ctx, const double * x, double * y, void *) { static_cast<const Operator *>(ctx)->action(x, y); return 0; }
        inline int apply_float(void * ctx, const float * x, float * y, void *) { static_cast<const SinglePrecisionOperator *>(ctx)->action(x, y); return 0; }

        inline solver_out unpack(const cuddh_solver_out & so, const std::vector<double> & res, const std::vector<double> & tim)
        {
            solver_out out;
            out.success = so.success != 0;
            out.num_iter = so.num_iter;
            out.num_matvec = so.num_matvec;
            out.res_norm.assign(res.begin(), res.begin() + so.n_res);
            out.time.assign(tim.begin(), tim.begin() + so.n_res);
            return out;
        }
    } // namespace detail

    /// GMRES(m) for P A x = P b (left preconditioning)
    inline solver_out gmres(int n, double * x, const Operator * A, const double * b, const Operator * Precond, int m, int maxit,
                            double tol = 1e-6, int verbose = 0, double max_seconds = 6 * 60 * 60)
    {
        std::vector<double> res((size_t)maxit + 2), tim((size_t)maxit + 2);
        cuddh_solver_out so;
        cuddh_check(cuddh_b200_gmres_d(n, x, &detail::apply_double, const_cast<Operator *>(A), b, Precond ? &detail::apply_double : nullptr,
                                       const_cast<Operator *>(Precond), m, maxit, tol, verbose, max_seconds, &so, res.data(), tim.data(),
                                       maxit + 2, nullptr));
        return detail::unpack(so, res, tim);
    }

    /// GMRES(m) for A x = b
    inline solver_out gmres(int n, double * x, const Operator * A, const double * b, int m, int maxit, double tol = 1e-6, int verbose = 0,
                            double max_seconds = 6 * 60 * 60)
    {
        return gmres(n, x, A, b, nullptr, m, maxit, tol, verbose, max_seconds);
    }

    /// single-precision GMRES(m) for A x = b
    inline solver_out gmres(int n, float * x, const SinglePrecisionOperator * A, const float * b, int m, int maxit, float tol = 1e-4,
                            int verbose = 0, double max_seconds = 6 * 60 * 60)
    {
        std::vector<double> res((size_t)maxit + 2), tim((size_t)maxit + 2);
        cuddh_solver_out so;
        cuddh_check(cuddh_b200_gmres_f(n, x, &detail::apply_float, const_cast<SinglePrecisionOperator *>(A), b, m, maxit, tol, verbose,
                                       max_seconds, &so, res.data(), tim.data(), maxit + 2, nullptr));
        return detail::unpack(so, res, tim);
    }
} // namespace cuddh

#endif
