// Shared declarations of libcuddh_b200: error plumbing, device buffers, host-side discretisation
// objects (tables, mesh, H1 space, face space, assembly plan). Pure C++ here; CUDA lives in *.cu.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX 3: a no-op unless a profiler (nsys / ncu --nvtx) injects itself

namespace cb200
{
    // ---- errors: every C-ABI entry point catches and converts to a status + last_error string ----
    struct Error : std::runtime_error
    {
        int code;
        Error(int c, const std::string & m) : std::runtime_error(m), code(c) {}
    };
    void set_last_error(const std::string & msg);
    const char * get_last_error();

#define CB_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            throw cb200::Error((int)e__, std::string(#expr " failed: ") + cudaGetErrorString(e__));          \
    } while (0)

    extern std::atomic<int64_t> g_launches; // kernels launched by this library (bench.py's gpu_launches)
#define CB_LAUNCHED()                                                                                         \
    do {                                                                                                      \
        ++cb200::g_launches;                                                                                  \
        CB_CUDA(cudaGetLastError());                                                                          \
    } while (0)

#define CB_REQUIRE(cond, msg)                                                                                 \
    do {                                                                                                      \
        if (!(cond))                                                                                          \
            throw cb200::Error(-1, std::string(msg));                                                         \
    } while (0)

    // ---- NVTX ranges around the host-visible stages of the solve path (operator applies, DDH action / rhs / postprocess, GMRES
    // restart cycles, halo exchanges): `nsys profile` / `ncu --nvtx` then show the reference's call structure
    // (examples/Helmholtz.hpp:28-56, source/gmres.cpp:146-215, source/DDH.cpp:611-695) on the timeline
    struct NvtxRange
    {
        explicit NvtxRange(const char * name) { nvtxRangePushA(name); }
        ~NvtxRange() { nvtxRangePop(); }
        NvtxRange(const NvtxRange &) = delete;
        NvtxRange & operator=(const NvtxRange &) = delete;
    };

    // ---- host-side parallel loop for the embarrassingly parallel setup stages (elements / subdomains are independent);
    // chunks are contiguous index ranges, so every stage writes the same values wherever it ran: results do not depend on
    // the thread count. CUDDH_B200_SETUP_THREADS overrides the thread count (1 = serial).
    int setup_threads();
    template <class F>
    void parallel_for(int64_t n, F && body /* (int64_t begin, int64_t end, int thread) */)
    {
        const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(setup_threads(), n / 256));
        if (nt <= 1) {
            body((int64_t)0, n, 0);
            return;
        }
        std::vector<std::thread> th;
        const int64_t chunk = (n + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
            if (b >= e)
                break;
            th.emplace_back([&body, b, e, t] { body(b, e, t); });
        }
        for (auto & x : th)
            x.join();
    }

    // ---- device buffer (64-bit sizes; the reference's int byte counts overflow at 2 GiB) ----
    template <typename T>
    struct DevBuf
    {
        T * p = nullptr;
        size_t n = 0;
        DevBuf() = default;
        explicit DevBuf(size_t n_) { alloc(n_); }
        DevBuf(const DevBuf &) = delete;
        DevBuf & operator=(const DevBuf &) = delete;
        DevBuf(DevBuf && o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
        DevBuf & operator=(DevBuf && o) noexcept
        {
            if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
            return *this;
        }
        ~DevBuf() { release(); }
        void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
        void alloc(size_t n_)
        {
            release();
            n = n_;
            if (n) CB_CUDA(cudaMalloc((void **)&p, n * sizeof(T)));
        }
        void zero(cudaStream_t s = 0) { if (n) CB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
        void upload(const T * h, size_t cnt)
        {
            if (cnt != n) alloc(cnt);
            if (cnt) CB_CUDA(cudaMemcpy(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice));
        }
        void upload(const std::vector<T> & h) { upload(h.data(), h.size()); }
        std::vector<T> download() const
        {
            std::vector<T> h(n);
            if (n) CB_CUDA(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
            return h;
        }
    };

    // ---- 1-D tables: reference source/QuadratureRule.cpp, source/Basis.cpp ----
    enum QuadType { GAUSS_LEGENDRE = 0, GAUSS_LOBATTO = 1 };
    void quadrature_rule(int n, int type, double * x, double * w);

    struct Basis
    {
        int n;
        std::vector<double> x, w;  // GLL nodes / weights
        std::vector<double> wb;    // barycentric weights
        explicit Basis(int n);
        void eval(int m, const double * xq, double * P /* (m,n) col-major */) const;
        void deriv(int m, const double * xq, double * D /* (m,n) col-major */) const;
    };

    // ---- mesh: reference source/Mesh2D.cpp ----
    struct Mesh
    {
        int64_t n_nodes = 0, n_elem = 0, n_edges = 0;
        int nx = 0, ny = 0;                  // > 0 iff built by uniform_rect (structured tiling hint only)
        std::vector<double> xy;              // (2, n_nodes)
        std::vector<int> elems;              // (4, n_elem) CCW corners
        std::vector<int> edges;              // (8, n_edges): n0,n1,el0,el1,side0,side1,delta,is_boundary
        std::vector<int> elem_edges;         // (4, n_elem): edge id of local side s
        std::vector<int> boundary_edges, interior_edges;
        std::vector<double> edge_meas;       // StraightEdge::meas = length/2
        double min_h = 0, max_h = 0;

        static std::unique_ptr<Mesh> from_vertices(int64_t nv, const double * xy, int64_t nel, const int * elems);
        static std::unique_ptr<Mesh> uniform_rect(int nx, double ax, double bx, int ny, double ay, double by);
        void corners(int64_t el, double * c /* 8: x0,y0,x1,y1,.. */) const
        {
            for (int k = 0; k < 4; ++k) {
                const int v = elems[4 * el + k];
                c[2 * k] = xy[2 * (size_t)v];
                c[2 * k + 1] = xy[2 * (size_t)v + 1];
            }
        }
    };

    // ---- assembly plan: CTA patches, colours, patch-local numbering, shared-DOF slots ----
    struct PatchHdr
    {
        int elem_begin;   // first element slot (multiple of PE)
        int n_elem;       // real elements in this patch (<= PE)
        int pdof_begin;   // offset into gid[]
        int n_pdof;       // patch-local DOFs
        int n_int;        // first n_int are touched by this patch only (written straight to y)
        int slot_begin;   // offset into slot[] for the n_pdof - n_int shared ones
        int cptr_begin;   // offset into cptr[] (n_pdof + 1 entries): CSR patch-local DOF -> element-local entries
        int reserved;
    };

    struct Plan
    {
        int nb = 0, PE = 0;
        bool interior_affine = false;      // node-major plans: in every element the ids of the nodes strictly inside it are base + (i-1) + (j-1) * stride
                                           // (true for the first-touch numbering of H1Space on any mesh; the thread-pair kernel hands over 2 ids per element)
        bool node_major = false;           // L / cent laid out (PE, nb*nb) per patch instead of (nb*nb, PE): thread-per-element kernels
        int64_t n_patches = 0, n_slots_total = 0, n_shared = 0;
        int max_pdof = 0, max_nsh = 0;   // largest patch: DOFs, shared DOFs
        std::vector<PatchHdr> hdr;
        std::vector<int> gid;              // patch-local -> global DOF, per patch ascending inside each class
        std::vector<int> slot;             // shared patch-local DOF -> index into the partial buffer
        std::vector<uint16_t> L;           // (nb*nb, PE, n_patches) element-local node -> patch-local DOF
        std::vector<uint16_t> cptr;        // per patch n_pdof+1 offsets into its slice of cent
        std::vector<uint16_t> cent;        // (nb*nb*PE, n_patches): element-local entries (slot*nb*nb + node) grouped by DOF
        std::vector<int> slot_elem;        // (PE, n_patches) global element id of each slot, -1 = padding
        // node-major plans (volume_action_ws) list only the DOFs on element boundaries; the nodes strictly inside an element are
        // written by the thread that owns the element:
        std::vector<int> Ig;               // (4, PE, ceil(nb*nb/4), n_patches) global DOF of (node % 4, slot, node / 4), 0 = padding
        std::vector<uint16_t> cent4;       // node-major plans only: per patch-local DOF (indexed like gid) its first four CSR entries,
                                           // 0xFFFF = none; entry 3 == 0xFFFE: more than four, continue in cptr/cent from entry 3
        std::vector<int> target;           // node-major plans only: per patch-local DOF the global DOF (private) or partial slot (shared)
        std::vector<int> sh_gid, sh_ptr;   // shared DOFs: global id, CSR into the partial buffer (patch order)
        // device mirrors
        DevBuf<PatchHdr> d_hdr;
        DevBuf<int> d_gid, d_slot, d_slot_elem, d_sh_gid, d_sh_ptr, d_Ig;
        DevBuf<uint16_t> d_L, d_cptr, d_cent, d_cent4;
        DevBuf<int> d_target;
        bool on_device = false;
        void ensure_device();
    };

    // ---- H1 space: reference source/H1Space.cpp:11-127 ----
    struct H1Space
    {
        const Mesh * mesh;
        std::unique_ptr<Basis> basis;
        int nb;
        int64_t ndof = 0, n_elem = 0;
        std::vector<int> I;        // (nb, nb, n_elem)
        std::vector<double> xy;    // (2, ndof)
        DevBuf<int> d_I;
        DevBuf<double> d_xy, d_corners;  // d_corners: (8, n_elem)
        std::unique_ptr<Plan> plan;      // built lazily by the operators
        std::unique_ptr<Plan> plan_tpe;  // node-major plan with 32-multiple patches (thread-per-element kernels), lazy
        DevBuf<int> d_tr_ptr, d_tr_src;  // transposed map DOF -> element-local entries (ordered assembly), lazy
        void ensure_transpose();
        // every element a parallelogram (constant Jacobian)? true for Mesh2D::uniform_rect; decided once, on the host
        bool all_affine();
        int affine_state = -1;

        H1Space(const Mesh * mesh, int nb);
        const int * device_I();
        const double * device_xy();
        const double * device_corners();
        Plan & get_plan();       // builds (once) and uploads
        Plan & get_plan_host();  // builds (once), host arrays only
        Plan & get_plan_tpe();   // thread-per-element variant, builds (once) and uploads
    };

    // ---- face space: reference source/H1Space.cpp:129-219 ----
    struct FaceSpace
    {
        H1Space * fem;
        int nb;
        int64_t n_faces = 0, fdof = 0;
        std::vector<int> faces;    // global edge ids
        std::vector<int> I;        // (nb, n_faces) face-space index
        std::vector<int> proj;     // (fdof) H1 index
        DevBuf<int> d_I, d_proj;
        // DOF-centric incidence (deterministic face-mass assembly): CSR dof -> (face*nb + k)
        std::vector<int> inc_ptr, inc;
        DevBuf<int> d_inc_ptr, d_inc;
        DevBuf<double> d_meas;     // (n_faces) StraightEdge::measure
        DevBuf<int> d_faces;
        std::vector<double> h_meas;
        bool on_device = false;

        FaceSpace(H1Space * fem, int64_t nf, const int * faces);
        void ensure_device();
    };

    void build_plan(H1Space & fem, Plan & plan, bool tpe = false);
    void plan_self_check(H1Space & fem, bool tpe, int64_t stats[8]); // host-only, see h1space.cpp
} // namespace cb200
