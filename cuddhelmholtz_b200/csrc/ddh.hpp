// DDH: the FP32 substructured Helmholtz operator (path B of SURVEY §8a) — reference include/DDH.hpp,
// source/DDH.cpp, source/EnsembleSpace.cpp.
#pragma once
#include "common.hpp"

namespace cb200
{
    // Subdomain index maps with the reference's numbering (source/EnsembleSpace.cpp:11-287); host only.
    struct Ensemble
    {
        int n_spaces = 0, nb = 0;
        int mx_elems = 0, mx_faces = 0, mx_ndof = 0, mx_fdof = 0;
        int64_t n_shared = 0;
        std::vector<int> s_elems, s_faces, s_dof, s_fdof;
        std::vector<int> elems;      // (mx_elems, p)
        std::vector<int> faces;      // (mx_faces, p) global edge ids
        std::vector<int> face_side;  // (mx_faces, p)
        std::vector<int> sI;         // (nb, nb, mx_elems, p)
        std::vector<int> gI;         // (mx_ndof, p)
        std::vector<int> fI;         // (nb, mx_faces, p)
        std::vector<int> pI;         // (mx_fdof, p)
        std::vector<int> cmap;       // (4, n_shared)
        Ensemble(const H1Space & fem, int n_spaces, const int * labels);
    };

    struct DDH
    {
        H1Space * fem;
        int nb, block, nel1;         // nel1 = block / nb elements per subdomain side
        int n1;                      // unique DOFs per subdomain side = nel1*(nb-1)+1
        int64_t g_ndof, n_shared, n_lambda;
        int n_domains, nt, mx_dof, mx_fdof, mx_elem;
        double omega, dt;
        std::unique_ptr<Ensemble> en;

        // reference-layout host arrays (kept for parity tests / introspection)
        std::vector<int> B, gI, sI;
        std::vector<float> m, gmi, a, H, D, wh_filter, cs, sn, g;

        // device arrays in the kernel's grid layout: per subdomain, per unique DOF (Y*n1 + X)
        DevBuf<int> d_gid, d_bin, d_bout;
        DevBuf<float> d_a, d_m, d_pou, d_H, d_g, d_D, d_whf, d_cs, d_sn;
        // deterministic partition-of-unity assembly of postprocess(): per-DOF contribution slots
        DevBuf<int> d_asm_ptr, d_asm_src;
        DevBuf<double> d_contrib;
        // host copies of the grid-layout tables (uploaded lazily so that the index data can be built and
        // inspected without a GPU)
        std::vector<int> hg_gid, hg_bin, hg_bout, h_asm_ptr, h_asm_src;
        std::vector<float> hg_a, hg_m, hg_pou, hg_H;
        bool on_device = false;
        void ensure_device();
        // metric diagonal and identical in every element (uniform_rect): enables the register-tiled kernel
        bool reg_tiled_ok = false;

        DDH(double omega, const double * h_a, H1Space * fem, int nx, int ny, int block);
        void rhs(const double * f, float * b, cudaStream_t s);
        void action(const float * x, float * y, cudaStream_t s);
        void postprocess(const float * lambda, const double * f, double * u, cudaStream_t s);
        void apply_T_range(const float * x, float * t, int dom_begin, int dom_end, cudaStream_t s);
        void rhs_range(const double * f, float * b, int dom_begin, int dom_end, cudaStream_t s);
        void postprocess_range(const float * lambda, const double * f, double * u, int dom_begin, int dom_end, cudaStream_t s);
        void get_array(const char * name, void * out, int64_t cap_bytes, int64_t * count) const;
        double flops() const;

    private:
        void run(const double * x, double * y, const float * lambda, float * update, cudaStream_t s, int dom_begin = 0, int dom_end = -1);
    };
} // namespace cb200
