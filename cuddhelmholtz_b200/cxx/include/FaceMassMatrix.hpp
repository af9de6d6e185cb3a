// FaceMassMatrix: m(u, v) = <u, v> or <a(x) u, v> on a FaceSpace, and its lumped inverse
// (reference include/FaceMassMatrix.hpp). Vectors are FaceSpace vectors.
#ifndef CUDDH_FACE_MASS_MATRIX_HPP
#define CUDDH_FACE_MASS_MATRIX_HPP

#include "H1Space.hpp"
#include "HostDeviceArray.hpp"
#include "Operator.hpp"
#include "StiffnessMatrix.hpp"
#include "forall.hpp"
#include "linalg.hpp"

namespace cuddh
{
    class FaceMassMatrix : public Operator
    {
    public:
        FaceMassMatrix(const FaceSpace & fs_) : fs(fs_) { create(nullptr); }
        /// a: DEVICE, FaceSpace vector of the coefficient
        FaceMassMatrix(const double * a, const FaceSpace & fs_) : fs(fs_) { create(a); }

        void action(double c, const double * x, double * y) const override { op.apply(c, x, y); }
        void action(const double * x, double * y) const override { op.apply(x, y); }

    private:
        void create(const double * a)
        {
            cuddh_operator_t raw = nullptr;
            cuddh_check(cuddh_b200_facemass_create(fs.handle(), a, 0, &raw));
            op.reset(raw);
        }
        const FaceSpace & fs;
        detail::OperatorHandle op;
    };

    class DiagInvFaceMassMatrix : public Operator
    {
    public:
        DiagInvFaceMassMatrix(const FaceSpace & fs_) { create(nullptr, fs_); }
        DiagInvFaceMassMatrix(const double * a, const FaceSpace & fs_) { create(a, fs_); }

        void action(double c, const double * x, double * y) const override { op.apply(c, x, y); }
        void action(const double * x, double * y) const override { op.apply(x, y); }

    private:
        void create(const double * a, const FaceSpace & fs_)
        {
            cuddh_operator_t raw = nullptr;
            cuddh_check(cuddh_b200_diag_inv_facemass_create(fs_.handle(), a, &raw));
            op.reset(raw);
        }
        detail::OperatorHandle op;
    };
} // namespace cuddh

#endif
