// StiffnessMatrix: b(u, v) = (grad u, grad v), matrix-free (reference include/StiffnessMatrix.hpp:11-38). The action is
// the sm_100a patch kernel of csrc/operators.cu.
#ifndef CUDDH_STIFFNESS_MATRIX_HPP
#define CUDDH_STIFFNESS_MATRIX_HPP

#include <memory>

#include "H1Space.hpp"
#include "Operator.hpp"
#include "linalg.hpp"

namespace cuddh
{
    namespace detail
    {
        /// owner of a cuddh_operator_t with the Operator::action semantics
        class OperatorHandle
        {
        public:
            void reset(cuddh_operator_t raw) { h.reset(raw, [](cuddh_operator_t p) { cuddh_b200_operator_destroy(p); }); }
            void apply(double c, const double * x, double * y) const { cuddh_check(cuddh_b200_operator_apply(h.get(), c, 1, x, y, nullptr)); }
            void apply(const double * x, double * y) const { cuddh_check(cuddh_b200_operator_apply(h.get(), 1.0, 0, x, y, nullptr)); }
            cuddh_operator_t get() const { return h.get(); }

        private:
            std::shared_ptr<cuddh_operator_s> h;
        };
    } // namespace detail

    class StiffnessMatrix : public Operator
    {
    public:
        /// Gauss-Legendre rule with n_basis + 1 points
        StiffnessMatrix(const H1Space & fem_) : fem(fem_)
        {
            cuddh_operator_t raw = nullptr;
            cuddh_check(cuddh_b200_stiffness_create(fem.handle(), 0, CUDDH_GAUSS_LEGENDRE, &raw));
            op.reset(raw);
        }
        StiffnessMatrix(const H1Space & fem_, const QuadratureRule & quad) : fem(fem_)
        {
            cuddh_operator_t raw = nullptr;
            cuddh_check(cuddh_b200_stiffness_create(fem.handle(), quad.size(),
                                                    quad.type() == QuadratureRule::GaussLegendre ? CUDDH_GAUSS_LEGENDRE : CUDDH_GAUSS_LOBATTO, &raw));
            op.reset(raw);
        }
        ~StiffnessMatrix() = default;

        /// y[i] <- y[i] + c * (grad x, grad phi[i])
        void action(double c, const double * x, double * y) const override { op.apply(c, x, y); }
        /// y[i] <- (grad x, grad phi[i])
        void action(const double * x, double * y) const override { op.apply(x, y); }

    private:
        const H1Space & fem;
        detail::OperatorHandle op;
    };
} // namespace cuddh

#endif
