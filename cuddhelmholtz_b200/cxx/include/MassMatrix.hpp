// MassMatrix: m(u, v) = (u, v) or (a(x) u, v); DiagInvMassMatrix: lumped (GLL-collocated) inverse
// (reference include/MassMatrix.hpp:14-68).
#ifndef CUDDH_MASS_MATRIX_HPP
#define CUDDH_MASS_MATRIX_HPP

#include "Basis.hpp"
#include "H1Space.hpp"
#include "HostDeviceArray.hpp"
#include "Mesh2D.hpp"
#include "StiffnessMatrix.hpp"
#include "forall.hpp"
#include "linalg.hpp"

namespace cuddh
{
    class MassMatrix : public Operator
    {
    public:
        MassMatrix(const H1Space & fem_) : fem(fem_) { create(nullptr); }
        /// a: DEVICE, H1Space vector of the coefficient
        MassMatrix(const double * a, const H1Space & fem_) : fem(fem_) { create(a); }

        void action(double c, const double * x, double * y) const override { op.apply(c, x, y); }
        void action(const double * x, double * y) const override { op.apply(x, y); }

    private:
        void create(const double * a)
        {
            cuddh_operator_t raw = nullptr;
            cuddh_check(cuddh_b200_mass_create(fem.handle(), a, 0, &raw));
            op.reset(raw);
        }
        const H1Space & fem;
        detail::OperatorHandle op;
    };

    class DiagInvMassMatrix : public Operator
    {
    public:
        DiagInvMassMatrix(const H1Space & fem_) : fem(fem_) { create(nullptr); }
        DiagInvMassMatrix(const double * a, const H1Space & fem_) : fem(fem_) { create(a); }

        void action(double c, const double * x, double * y) const override { op.apply(c, x, y); }
        void action(const double * x, double * y) const override { op.apply(x, y); }

    private:
        void create(const double * a)
        {
            cuddh_operator_t raw = nullptr;
            cuddh_check(cuddh_b200_diag_inv_mass_create(fem.handle(), a, &raw));
            op.reset(raw);
        }
        const H1Space & fem;
        detail::OperatorHandle op;
    };
} // namespace cuddh

#endif
