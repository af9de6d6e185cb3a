// see aux.cu
#pragma once
#include "common.hpp"

namespace cb200
{
    // which: 0 jacobians (2,2,nq,nq,nel), 1 measures (nq,nq,nel), 2 physical coordinates (2,nq,nq,nel); device output
    void element_metrics(H1Space * fem, int nq, const double * h_xq, int which, double * d_out, cudaStream_t s);
    // F += c * assemble( P^T g P ) ; h_P == null: g already holds nodal values (nb,nb,nel)
    void linear_functional_assemble(H1Space * fem, int nq, const double * h_P, const double * d_g, double c, double * d_F, cudaStream_t s);
    void face_linear_functional_assemble(FaceSpace * fs, int nq, const double * h_P, const double * d_g, double c, double * d_F, cudaStream_t s);
} // namespace cb200
