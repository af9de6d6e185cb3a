// Auxiliary device work behind the C++ header layer (SURVEY §8f "next" rows, needed by every example / test of the
// reference): element metric arrays on quadrature points and the deterministic assembly used by
// LinearFunctional / FaceLinearFunctional.
//
// Reference: Mesh2D::ElementMetricCollection (source/Mesh2D.cpp:173-227; serial host loop with virtual calls, then an
// H2D copy), lf_action / lf_fast (include/LinearFunctional.hpp:47-143), fl_action / fl_fast
// (include/FaceLinearFunctional.hpp:48-128). Here the metrics are evaluated on the device straight from the element
// corners, and the (f, phi) contraction is followed by an ordered gather through a transposed map instead of
// atomicAdd.
#include "aux.hpp"
#include <algorithm>

namespace cb200
{
    namespace
    {
        inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

        // which: 0 = jacobians (2,2,nq,nq,nel), 1 = measures (nq,nq,nel), 2 = physical coordinates (2,nq,nq,nel)
        __global__ void element_metrics_kernel(const int64_t nel, const int nq, const int which, const double * __restrict__ corners,
                                               const double * __restrict__ xq, double * __restrict__ out)
        {
            const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            const int nq2 = nq * nq;
            if (t >= nel * nq2)
                return;
            const int64_t el = t / nq2;
            const int rem = (int)(t - el * nq2);
            const int i = rem % nq, j = rem / nq;
            const double * c = corners + 8 * el;
            const double xi0 = xq[i], xi1 = xq[j];
            if (which == 2) { // source/Element.cpp:5-19
                const double b[4] = {0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1),
                                     0.25 * (1.0 + xi0) * (1.0 + xi1), 0.25 * (1.0 - xi0) * (1.0 + xi1)};
                double x0 = 0.0, x1 = 0.0;
                for (int k = 0; k < 4; ++k) {
                    x0 += c[2 * k] * b[k];
                    x1 += c[2 * k + 1] * b[k];
                }
                out[2 * t] = x0;
                out[2 * t + 1] = x1;
                return;
            }
            // source/Element.cpp:21-27
            const double J0 = 0.25 * ((1.0 - xi1) * (c[2] - c[0]) + (1.0 + xi1) * (c[4] - c[6]));
            const double J1 = 0.25 * ((1.0 - xi1) * (c[3] - c[1]) + (1.0 + xi1) * (c[5] - c[7]));
            const double J2 = 0.25 * ((1.0 - xi0) * (c[6] - c[0]) + (1.0 + xi0) * (c[4] - c[2]));
            const double J3 = 0.25 * ((1.0 - xi0) * (c[7] - c[1]) + (1.0 + xi0) * (c[5] - c[3]));
            if (which == 0) {
                out[4 * t] = J0;
                out[4 * t + 1] = J1;
                out[4 * t + 2] = J2;
                out[4 * t + 3] = J3;
            }
            else
                out[t] = J0 * J3 - J1 * J2;
        }

        // elemvec(tx,ty,el) = sum_i sum_j P(i,tx) P(j,ty) g(i,j,el)  (order of include/LinearFunctional.hpp:92-112)
        __global__ void lf_contract_kernel(const int64_t nel, const int nb, const int nq, const double * __restrict__ P,
                                           const double * __restrict__ g, double * __restrict__ ev)
        {
            const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            const int nb2 = nb * nb;
            if (t >= nel * nb2)
                return;
            const int64_t el = t / nb2;
            const int rem = (int)(t - el * nb2);
            const int tx = rem % nb, ty = rem / nb;
            const double * ge = g + (size_t)nq * nq * el;
            double qqu = 0.0;
            for (int i = 0; i < nq; ++i) {
                double qu = 0.0; // Pg[i][ty] = sum_j P(j,ty) g(i,j)
                for (int j = 0; j < nq; ++j)
                    qu += P[j + nq * ty] * ge[i + nq * j];
                qqu += P[i + nq * tx] * qu;
            }
            ev[t] = qqu;
        }

        // F[d] += c * sum over the DOF's element-local entries, ascending (deterministic replacement of atomicAdd)
        __global__ void gather_add_kernel(const int64_t n, const int * __restrict__ ptr, const int * __restrict__ src,
                                          const double * __restrict__ ev, const double c, double * __restrict__ F)
        {
            const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (d >= n)
                return;
            double s = 0.0;
            for (int k = ptr[d]; k < ptr[d + 1]; ++k)
                s += c * ev[src[k]];
            F[d] += s;
        }

        // face version: F[d] += c * sum_{(f,k) incident to d} sum_i P(i,k) g(i,f)   (P == null: g is (nb, nf) already)
        __global__ void face_lf_kernel(const int64_t fdof, const int nb, const int nq, const double * __restrict__ P,
                                       const double * __restrict__ g, const int * __restrict__ inc_ptr, const int * __restrict__ inc,
                                       const double c, double * __restrict__ F)
        {
            const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
            if (d >= fdof)
                return;
            double s = 0.0;
            for (int t = inc_ptr[d]; t < inc_ptr[d + 1]; ++t) {
                const int fk = inc[t];
                const int f = fk / nb, k = fk - f * nb;
                double v = 0.0;
                if (P) {
                    for (int i = 0; i < nq; ++i)
                        v += P[i + nq * k] * g[i + (size_t)nq * f];
                }
                else
                    v = g[k + (size_t)nb * f];
                s += c * v;
            }
            F[d] += s;
        }
    } // namespace

    void element_metrics(H1Space * fem, int nq, const double * h_xq, int which, double * d_out, cudaStream_t s)
    {
        DevBuf<double> d_x;
        d_x.upload(h_xq, (size_t)nq);
        const int64_t n = fem->n_elem * (int64_t)nq * nq;
        if (n == 0)
            return;
        element_metrics_kernel<<<nblk(n), 256, 0, s>>>(fem->n_elem, nq, which, fem->device_corners(), d_x.p, d_out);
        CB_LAUNCHED();
        CB_CUDA(cudaStreamSynchronize(s)); // d_x is freed on return
    }

    void H1Space::ensure_transpose()
    {
        if (d_tr_ptr.p)
            return;
        const int64_t N = (int64_t)nb * nb * n_elem;
        std::vector<int> ptr((size_t)ndof + 1, 0), src((size_t)N);
        for (int64_t t = 0; t < N; ++t)
            ptr[I[t] + 1]++;
        for (int64_t d = 0; d < ndof; ++d)
            ptr[d + 1] += ptr[d];
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (int64_t t = 0; t < N; ++t)
            src[cur[I[t]]++] = (int)t;
        d_tr_ptr.upload(ptr);
        d_tr_src.upload(src);
    }

    void linear_functional_assemble(H1Space * fem, int nq, const double * h_P, const double * d_g, double c, double * d_F, cudaStream_t s)
    {
        fem->ensure_transpose();
        const int nb = fem->nb;
        const int64_t N = (int64_t)nb * nb * fem->n_elem;
        const double * ev = d_g;
        DevBuf<double> tmp, d_P;
        if (h_P) {
            d_P.upload(h_P, (size_t)nq * nb);
            tmp.alloc((size_t)N);
            lf_contract_kernel<<<nblk(N), 256, 0, s>>>(fem->n_elem, nb, nq, d_P.p, d_g, tmp.p);
            CB_LAUNCHED();
            ev = tmp.p;
        }
        gather_add_kernel<<<nblk(fem->ndof), 256, 0, s>>>(fem->ndof, fem->d_tr_ptr.p, fem->d_tr_src.p, ev, c, d_F);
        CB_LAUNCHED();
        CB_CUDA(cudaStreamSynchronize(s)); // temporaries are freed on return
    }

    void face_linear_functional_assemble(FaceSpace * fs, int nq, const double * h_P, const double * d_g, double c, double * d_F, cudaStream_t s)
    {
        if (fs->fdof == 0)
            return;
        fs->ensure_device();
        DevBuf<double> d_P;
        if (h_P)
            d_P.upload(h_P, (size_t)nq * fs->nb);
        face_lf_kernel<<<nblk(fs->fdof, 128), 128, 0, s>>>(fs->fdof, fs->nb, nq, h_P ? d_P.p : nullptr, d_g, fs->d_inc_ptr.p, fs->d_inc.p, c, d_F);
        CB_LAUNCHED();
        CB_CUDA(cudaStreamSynchronize(s));
    }
} // namespace cb200
