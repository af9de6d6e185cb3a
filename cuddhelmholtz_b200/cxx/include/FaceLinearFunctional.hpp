// FaceLinearFunctional: F[i] += c * <f, phi[i]> over the faces of a FaceSpace, f a __device__ callable
// (reference include/FaceLinearFunctional.hpp). Same split as LinearFunctional: the user's lambda fills
// w_i * measure * f(x_i) per face point, the library contracts and assembles in a fixed order.
#ifndef CUDDH_FACE_LINEAR_FUNCTIONAL_HPP
#define CUDDH_FACE_LINEAR_FUNCTIONAL_HPP

#include "H1Space.hpp"
#include "HostDeviceArray.hpp"
#include "forall.hpp"
#include "linalg.hpp"

namespace cuddh
{
    class FaceLinearFunctional
    {
    public:
        FaceLinearFunctional(const FaceSpace & fs_)
            : fs(fs_), metrics(fs_.metrics(fs_.h1_space().basis().quadrature())), fdof(fs_.size()), n_faces(fs_.n_faces()),
              n_basis(fs_.h1_space().basis().size()), n_quad(n_basis), fast(true), _w(n_quad)
        {
            load_weights(fs.h1_space().basis().quadrature());
        }

        FaceLinearFunctional(const FaceSpace & fs_, const QuadratureRule & quad)
            : fs(fs_), metrics(fs_.metrics(quad)), fdof(fs_.size()), n_faces(fs_.n_faces()), n_basis(fs_.h1_space().basis().size()),
              n_quad(quad.size()), fast(false), _w(n_quad), _P(n_quad * n_basis)
        {
            load_weights(quad);
            fs.h1_space().basis().eval(n_quad, quad.x(), _P.host_write());
        }

        /// F[i] <- F[i] + c * <f, phi[i]>
        template <typename Func>
        void action(double c, Func && f, double * F) const
        {
            const int nq = n_quad, npts = n_quad * n_faces;
            if (npts == 0)
                return;
            _g.resize(npts);
            double * g = _g.device_write();
            const double * wq = _w.device_read();
            const double * detJ = metrics.measures(MemorySpace::DEVICE);          // (nq, n_faces)
            const double * X = metrics.physical_coordinates(MemorySpace::DEVICE); // (2, nq, n_faces)
            forall(npts, [=] __device__(int t) -> void {
                double xi[2] = {X[2 * t], X[2 * t + 1]};
                g[t] = f(xi) * wq[t % nq] * detJ[t];
            });
            cuddh_check(cuddh_b200_face_linear_functional_assemble(fs.handle(), nq, fast ? nullptr : _P.host_read(), g, c, F, nullptr));
        }

        /// F[i] <- <f, phi[i]>
        template <typename Func>
        void action(Func && f, double * F) const
        {
            zeros(fdof, F);
            action(1.0, f, F);
        }

    private:
        void load_weights(const QuadratureRule & q)
        {
            double * hw = _w.host_write();
            for (int i = 0; i < n_quad; ++i)
                hw[i] = q.w(i);
        }

        const FaceSpace & fs;
        const Mesh2D::EdgeMetricCollection & metrics;
        const int fdof, n_faces, n_basis, n_quad;
        const bool fast;
        host_device_dvec _w, _P;
        mutable host_device_dvec _g;
    };
} // namespace cuddh

#endif
