// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// ref_driver: a small command-line program, written for this repo, that links the
// UNMODIFIED reference library (the sources are compiled where they lie under
// /root/reference by oracle/Makefile; nothing is copied) and calls its PUBLIC API
// to dump index maps, tables and operator / solver outputs into a flat binary
// container ("RDMP") that tests/ and bench.py read with numpy.
//
// Host-only commands (tables, h1, ensemble) run in the build container and generate
// the committed fixtures under tests/golden/ (see scripts/make_golden.py).
// Device commands (ops, helm_gmres, ddh, time_ops, time_ddh) need a GPU: on the B200
// box they make the reference's own kernels (sm_100 build) the parity oracle and
// the "reference GPU" timing column of bench.py.
//
// Only DDH's private index arrays are reached through `#define private public`
// (DDH has no accessor for them); everything else goes through public methods.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <random>
#include <fstream>
#include <chrono>

#define private public
#include "include/DDH.hpp"
#undef private
#include "cuddh.hpp"
#include "examples/Helmholtz.hpp"

using namespace cuddh;

// ----------------------------------------------------------------------------------
// RDMP container: repeated records { char name[48]; int32 dtype; int64 count; data }.
// dtype: 0 = int32, 1 = float64, 2 = float32
// ----------------------------------------------------------------------------------
struct Dump
{
    FILE * f;
    explicit Dump(const std::string & path)
    {
        f = fopen(path.c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    }
    ~Dump() { fclose(f); }

    void put(const char * name, int dtype, int64_t count, const void * data, size_t elsize)
    {
        char nm[48];
        memset(nm, 0, sizeof nm);
        strncpy(nm, name, 47);
        int32_t dt = dtype;
        fwrite(nm, 1, 48, f);
        fwrite(&dt, 4, 1, f);
        fwrite(&count, 8, 1, f);
        if (count > 0) fwrite(data, elsize, (size_t)count, f);
    }
    void ints(const char * name, int64_t n, const int * p) { put(name, 0, n, p, 4); }
    void dbls(const char * name, int64_t n, const double * p) { put(name, 1, n, p, 8); }
    void flts(const char * name, int64_t n, const float * p) { put(name, 2, n, p, 4); }
    void i1(const char * name, int v) { ints(name, 1, &v); }
    void d1(const char * name, double v) { dbls(name, 1, &v); }
};

static Mesh2D load_mesh(const std::string & spec)
{
    // "rect:<nx>"  or  "file:<path>" (text: "nv nel", then nv lines "x y", then nel lines "a b c d")
    if (spec.rfind("rect:", 0) == 0)
    {
        int nx = atoi(spec.c_str() + 5);
        return Mesh2D::uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0);
    }
    else if (spec.rfind("file:", 0) == 0)
    {
        std::ifstream in(spec.substr(5));
        if (!in) { fprintf(stderr, "cannot open mesh %s\n", spec.c_str()); exit(2); }
        int nv, nel;
        in >> nv >> nel;
        std::vector<double> x(2 * (size_t)nv);
        std::vector<int> e(4 * (size_t)nel);
        for (int i = 0; i < nv; ++i) in >> x[2*i] >> x[2*i+1];
        for (int i = 0; i < nel; ++i) in >> e[4*i] >> e[4*i+1] >> e[4*i+2] >> e[4*i+3];
        return Mesh2D::from_vertices(nv, x.data(), nel, e.data());
    }
    fprintf(stderr, "bad mesh spec %s\n", spec.c_str());
    exit(2);
}

template <typename T>
static std::vector<T> d2h(const T * d, size_t n)
{
    std::vector<T> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
}

static void fill_uniform(std::vector<double> & v, uint64_t seed)
{
    std::mt19937_64 gen(seed);
    for (auto & x : v)
        x = 2.0 * ((gen() >> 11) * (1.0 / 9007199254740992.0)) - 1.0; // uniform(-1,1), 53 bits
}

// ----------------------------------------------------------------------------------
// host-only commands
// ----------------------------------------------------------------------------------
static int cmd_tables(int nb, int nq, const std::string & out)
{
    Dump d(out);
    Basis basis(nb);
    auto & gll = basis.quadrature();
    d.dbls("gll_x", nb, gll.x().data());
    d.dbls("gll_w", nb, gll.w().data());
    d.dbls("Dnodes", nb * nb, basis.derivative_matrix().data());
    d.dbls("Mbasis", nb * nb, basis.mass_matrix().data());
    QuadratureRule gl(nq, QuadratureRule::GaussLegendre);
    d.dbls("gl_x", nq, gl.x().data());
    d.dbls("gl_w", nq, gl.w().data());
    std::vector<double> P(nq * nb), D(nq * nb);
    basis.eval(nq, gl.x(), P.data());
    basis.deriv(nq, gl.x(), D.data());
    d.dbls("P", nq * nb, P.data());
    d.dbls("D", nq * nb, D.data());
    QuadratureRule gll2(nq, QuadratureRule::GaussLobatto);
    d.dbls("gll_nq_x", nq, gll2.x().data());
    d.dbls("gll_nq_w", nq, gll2.w().data());
    return 0;
}

static void dump_mesh(Dump & d, const Mesh2D & mesh)
{
    const int ne = mesh.n_edges();
    std::vector<int> E(8 * (size_t)ne); // nodes[2], elements[2], sides[2], delta, type
    for (int e = 0; e < ne; ++e)
    {
        const Edge * edge = mesh.edge(e);
        const bool interior = edge->type == FaceType::INTERIOR;
        E[8*e+0] = edge->nodes[0];
        E[8*e+1] = edge->nodes[1];
        E[8*e+2] = edge->elements[0];
        E[8*e+3] = interior ? edge->elements[1] : -1;
        E[8*e+4] = edge->sides[0];
        E[8*e+5] = interior ? edge->sides[1] : -1;
        E[8*e+6] = edge->delta;
        E[8*e+7] = interior ? 0 : 1;
    }
    d.i1("n_elem", mesh.n_elem());
    d.i1("n_nodes", mesh.n_nodes());
    d.i1("n_edges", ne);
    d.ints("edges", 8 * (int64_t)ne, E.data());
    ivec b = mesh.boundary_edges();
    d.ints("boundary_edges", b.size(), b.data());
    d.d1("min_h", mesh.min_h());
    d.d1("max_h", mesh.max_h());
}

static int cmd_h1(const std::string & meshspec, int nb, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = load_mesh(meshspec);
    dump_mesh(d, mesh);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    d.i1("ndof", ndof);
    auto I = fem.global_indices(MemorySpace::HOST);
    d.ints("I", (int64_t)nb * nb * mesh.n_elem(), I.data());
    auto xy = fem.physical_coordinates(MemorySpace::HOST);
    d.dbls("xy", 2 * (int64_t)ndof, xy.data());

    ivec bf = mesh.boundary_edges();
    FaceSpace fs(fem, bf.size(), bf.data());
    d.i1("fdof", fs.size());
    d.ints("face_I", (int64_t)nb * fs.n_faces(), fs.subspace_indices(MemorySpace::HOST).data());
    d.ints("face_proj", fs.size(), fs.global_indices(MemorySpace::HOST).data());
    return 0;
}

static std::vector<int> ddh_labels(int nx, int ny, int nb, int block, int & n_domains)
{
    const int epd = block / nb;
    const int ndx = nx / epd, ndy = ny / epd;
    n_domains = ndx * ndy;
    std::vector<int> lab((size_t)nx * ny);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            lab[i + (size_t)nx * j] = (i / epd) + ndx * (j / epd);
    return lab;
}

static int cmd_ensemble(int nx, int nb, int block, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = Mesh2D::uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    int n_domains;
    auto lab = ddh_labels(nx, nx, nb, block, n_domains);
    EnsembleSpace ef(fem, n_domains, lab.data());
    d.i1("n_domains", n_domains);
    d.i1("ndof", fem.size());
    auto sizes = ef.sizes(MemorySpace::HOST);
    auto fsizes = ef.fsizes(MemorySpace::HOST);
    auto nel = ef.n_elems(MemorySpace::HOST);
    auto nfc = ef.n_faces(MemorySpace::HOST);
    d.ints("sizes", n_domains, sizes.data());
    d.ints("fsizes", n_domains, fsizes.data());
    d.ints("n_elems", n_domains, nel.data());
    d.ints("n_faces", n_domains, nfc.data());
    auto gI = ef.global_indices(MemorySpace::HOST);
    d.ints("gI", gI.size(), gI.data());
    auto el = ef.elements(MemorySpace::HOST);
    d.ints("elements", el.size(), el.data());
    auto fc = ef.faces(MemorySpace::HOST);
    d.ints("faces", fc.size(), fc.data());
    auto sI = ef.subspace_indices(MemorySpace::HOST);
    d.ints("sI", sI.size(), sI.data());
    auto fI = ef.face_indices(MemorySpace::HOST);
    d.ints("fI", fI.size(), fI.data());
    auto pI = ef.face_proj(MemorySpace::HOST);
    d.ints("pI", pI.size(), pI.data());
    auto cm = ef.connectivity_map(MemorySpace::HOST);
    d.ints("cmap", cm.size(), cm.data());
    return 0;
}

// ----------------------------------------------------------------------------------
// device commands
// ----------------------------------------------------------------------------------
__device__ static double f_stiff(const double X[2])
{
    const double x = X[0], y = X[1];
    double x5 = std::pow(x, 5);
    double y3 = std::pow(y, 3);
    return (x5 - 5.0 * x) * (y3 - 3.0 * y);
}

__device__ static double f_mass(const double X[2])
{
    const double x = X[0], y = X[1];
    return 3.0 * x * x - 2.0 * x * y + y + 1.0;
}

__device__ static double f_coef(const double X[2])
{
    return 1.0 + 0.5 * std::sin(M_PI * X[0]) * std::cos(M_PI * X[1]);
}

// "lite": only what the large steady-state parity cases need (a2, af, the seeded [u; v], S u, S accumulate, weighted M u,
// M accumulate, Helmholtz composite) - one third of the bytes of the full dump
static int cmd_ops_lite(const std::string & meshspec, int nb, double omega, uint64_t seed, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = load_mesh(meshspec);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    d.i1("ndof", ndof);
    d.i1("nb", nb);
    ivec bf = mesh.boundary_edges();
    FaceSpace fs(fem, bf.size(), bf.data());
    const int fdof = fs.size();
    d.i1("fdof", fdof);
    host_device_dvec _a2(ndof), _af(fdof), _y(ndof), _u(2 * ndof), _Au(2 * ndof);
    double * a2 = _a2.device_write();
    double * af = _af.device_write();
    double * y = _y.device_write();
    double * u = _u.device_write();
    double * Au = _Au.device_write();
    auto X = fem.physical_coordinates(MemorySpace::DEVICE);
    forall(ndof, [=] __device__ (int i) -> void
    {
        const double xi[] = {X(0, i), X(1, i)};
        const double c = f_coef(xi);
        a2[i] = c * c;
    });
    {
        auto proj = fs.global_indices(MemorySpace::DEVICE);
        forall(fdof, [=] __device__ (int i) -> void
        {
            const int g = proj(i);
            const double xi[] = {X(0, g), X(1, g)};
            af[i] = f_coef(xi);
        });
    }
    std::vector<double> hu(2 * (size_t)ndof);
    fill_uniform(hu, seed);
    cudaMemcpy(u, hu.data(), hu.size() * sizeof(double), cudaMemcpyHostToDevice);
    d.dbls("a2", ndof, d2h(a2, ndof).data());
    d.dbls("af", fdof, d2h(af, fdof).data());
    d.dbls("helm_x", 2 * (int64_t)ndof, hu.data());
    {
        StiffnessMatrix S(fem);
        S.action(u, y);
        d.dbls("S_u", ndof, d2h(y, ndof).data());
        S.action(-0.75, u + ndof, y);
        d.dbls("S_acc", ndof, d2h(y, ndof).data());
    }
    {
        MassMatrix Mw(a2, fem);
        Mw.action(u, y);
        d.dbls("Mw_u", ndof, d2h(y, ndof).data());
        Mw.action(2.5, u + ndof, y);
        d.dbls("Mw_acc", ndof, d2h(y, ndof).data());
    }
    {
        Helmholtz A(omega, a2, af, fem, fs);
        A.action(u, Au);
        d.d1("omega", omega);
        d.dbls("helm_Ax", 2 * (int64_t)ndof, d2h(Au, 2 * (size_t)ndof).data());
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    return 0;
}

// Mesh2D::ElementMetricCollection (source/Mesh2D.cpp:173-227) on the Gauss-Legendre rule with nq points, LinearFunctional::action
// (include/LinearFunctional.hpp:145-181; collocated and nq-point rule) and FaceLinearFunctional::action on the boundary face
// space (include/FaceLinearFunctional.hpp:130-164), for fixed polynomial / smooth integrands
static int cmd_metrics_lf(const std::string & meshspec, int nb, int nq, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = load_mesh(meshspec);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    const int nel = mesh.n_elem();
    d.i1("ndof", ndof);
    d.i1("n_elem", nel);
    QuadratureRule q(nq, QuadratureRule::GaussLegendre);
    d.dbls("xq", nq, q.x().data());
    auto & em = mesh.element_metrics(q);
    d.dbls("J", 4 * (int64_t)nq * nq * nel, d2h(em.jacobians(MemorySpace::DEVICE), 4 * (size_t)nq * nq * nel).data());
    d.dbls("detJ", (int64_t)nq * nq * nel, d2h(em.measures(MemorySpace::DEVICE), (size_t)nq * nq * nel).data());
    d.dbls("xphys", 2 * (int64_t)nq * nq * nel, d2h(em.physical_coordinates(MemorySpace::DEVICE), 2 * (size_t)nq * nq * nel).data());

    host_device_dvec _F(ndof);
    double * F = _F.device_write();
    {
        LinearFunctional l(fem);
        l.action([] __device__ (const double X[2]) -> double {return f_mass(X);}, F);
        d.dbls("lf_fast", ndof, d2h(F, ndof).data());
        LinearFunctional l2(fem, q);
        l2.action([] __device__ (const double X[2]) -> double {return f_mass(X);}, F);
        d.dbls("lf_quad", ndof, d2h(F, ndof).data());
        l2.action(-0.5, [] __device__ (const double X[2]) -> double {return f_coef(X);}, F);
        d.dbls("lf_quad_acc", ndof, d2h(F, ndof).data());
    }
    {
        ivec bf = mesh.boundary_edges();
        FaceSpace fs(fem, bf.size(), bf.data());
        const int fdof = fs.size();
        d.i1("fdof", fdof);
        host_device_dvec _G(fdof);
        double * G = _G.device_write();
        FaceLinearFunctional fl(fs);
        fl.action([] __device__ (const double X[2]) -> double {return f_mass(X);}, G);
        d.dbls("fl_fast", fdof, d2h(G, fdof).data());
        FaceLinearFunctional fl2(fs, q);
        fl2.action([] __device__ (const double X[2]) -> double {return f_mass(X);}, G);
        d.dbls("fl_quad", fdof, d2h(G, fdof).data());
        fl2.action(1.5, [] __device__ (const double X[2]) -> double {return f_coef(X);}, G);
        d.dbls("fl_quad_acc", fdof, d2h(G, fdof).data());
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    return 0;
}

static int cmd_ops(const std::string & meshspec, int nb, double omega, uint64_t seed, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = load_mesh(meshspec);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    d.i1("ndof", ndof);
    d.i1("nb", nb);

    ivec bf = mesh.boundary_edges();
    FaceSpace fs(fem, bf.size(), bf.data());
    const int fdof = fs.size();
    d.i1("fdof", fdof);

    host_device_dvec _fs(ndof), _fm(ndof), _a2(ndof), _af(fdof), _y(ndof), _xf(fdof), _yf(fdof);
    double * fsv = _fs.device_write();
    double * fmv = _fm.device_write();
    double * a2 = _a2.device_write();
    double * af = _af.device_write();
    double * y = _y.device_write();
    double * xf = _xf.device_write();
    double * yf = _yf.device_write();

    auto X = fem.physical_coordinates(MemorySpace::DEVICE);
    forall(ndof, [=] __device__ (int i) -> void
    {
        const double xi[] = {X(0, i), X(1, i)};
        fsv[i] = f_stiff(xi);
        fmv[i] = f_mass(xi);
        const double c = f_coef(xi);
        a2[i] = c * c;
    });
    // a on the face space = restriction of the nodal coefficient
    {
        auto proj = fs.global_indices(MemorySpace::DEVICE);
        forall(fdof, [=] __device__ (int i) -> void
        {
            const int g = proj(i);
            const double xi[] = {X(0, g), X(1, g)};
            af[i] = f_coef(xi);
        });
    }
    d.dbls("f_stiff", ndof, d2h(fsv, ndof).data());
    d.dbls("f_mass", ndof, d2h(fmv, ndof).data());
    d.dbls("a2", ndof, d2h(a2, ndof).data());
    d.dbls("af", fdof, d2h(af, fdof).data());

    {   // stiffness, default quadrature (nb+1) and the tests' nb+2
        StiffnessMatrix S(fem);
        S.action(fsv, y);
        d.dbls("S_default_f", ndof, d2h(y, ndof).data());
        S.action(-0.75, fmv, y); // accumulate form
        d.dbls("S_default_acc", ndof, d2h(y, ndof).data());
        QuadratureRule q(nb + 2, QuadratureRule::GaussLegendre);
        StiffnessMatrix S2(fem, q);
        S2.action(fsv, y);
        d.dbls("S_q2_f", ndof, d2h(y, ndof).data());
    }
    {   // mass, unweighted and weighted
        MassMatrix M(fem);
        M.action(fmv, y);
        d.dbls("M_f", ndof, d2h(y, ndof).data());
        MassMatrix Mw(a2, fem);
        Mw.action(fmv, y);
        d.dbls("Mw_f", ndof, d2h(y, ndof).data());
        Mw.action(2.5, fsv, y);
        d.dbls("Mw_acc", ndof, d2h(y, ndof).data());
        DiagInvMassMatrix Mi(fem);
        Mi.action(fmv, y);
        d.dbls("Mi_f", ndof, d2h(y, ndof).data());
        DiagInvMassMatrix Miw(a2, fem);
        Miw.action(fmv, y);
        d.dbls("Miw_f", ndof, d2h(y, ndof).data());
    }
    {   // face mass on the boundary face space
        fs.restrict(fmv, xf);
        d.dbls("restrict_f", fdof, d2h(xf, fdof).data());
        FaceMassMatrix H(fs);
        H.action(xf, yf);
        d.dbls("H_f", fdof, d2h(yf, fdof).data());
        FaceMassMatrix Hw(af, fs);
        Hw.action(xf, yf);
        d.dbls("Hw_f", fdof, d2h(yf, fdof).data());
        DiagInvFaceMassMatrix Hi(fs);
        Hi.action(xf, yf);
        d.dbls("Hi_f", fdof, d2h(yf, fdof).data());
        cudaMemcpy(y, fsv, ndof * sizeof(double), cudaMemcpyDeviceToDevice);
        fs.prolong(yf, y);
        d.dbls("prolong", ndof, d2h(y, ndof).data());
        fs.orth(y);
        d.dbls("orth", ndof, d2h(y, ndof).data());
    }
    {   // the Helmholtz composite of examples/Helmholtz.hpp on a seeded random [u;v]
        std::vector<double> hu(2 * (size_t)ndof);
        fill_uniform(hu, seed);
        host_device_dvec _u(2 * ndof), _Au(2 * ndof);
        double * u = _u.device_write();
        double * Au = _Au.device_write();
        cudaMemcpy(u, hu.data(), hu.size() * sizeof(double), cudaMemcpyHostToDevice);
        Helmholtz A(omega, a2, af, fem, fs);
        A.action(u, Au);
        d.d1("omega", omega);
        d.dbls("helm_x", 2 * (int64_t)ndof, hu.data());
        d.dbls("helm_Ax", 2 * (int64_t)ndof, d2h(Au, 2 * (size_t)ndof).data());
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    return 0;
}

__device__ static double f_src(const double X[2], double omega)
{
    const double x = X[0], y = X[1];
    double s = omega * omega;
    double r = (x+0.5)*(x+0.5) + y * y;
    double F = s / M_PI * std::exp(-s * r);
    r = (x-0.5)*(x-0.5) + (y+0.5)*(y+0.5);
    F += s / M_PI * std::exp(-s * r);
    return F;
}

__device__ static double f_alpha(const double X[2])
{
    const double r = X[0]*X[0] + X[1]*X[1];
    return (r < 0.0625) ? 0.2 : 1.0;
}

// FP64 Helmholtz GMRES on any mesh: nodal coefficient a(x) (interpolated, not projected)
static int cmd_helm_gmres(const std::string & meshspec, int nb, double omega, int m, int maxit, double tol, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = load_mesh(meshspec);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    ivec bf = mesh.boundary_edges();
    FaceSpace fs(fem, bf.size(), bf.data());
    const int fdof = fs.size();
    const int N = 2 * ndof;

    host_device_dvec U(N), b(N), a2x(ndof), ax(fdof);
    double * d_U = U.device_write();
    double * d_b = b.device_write();
    double * d_a2 = a2x.device_write();
    double * d_a = ax.device_write();

    auto X = fem.physical_coordinates(MemorySpace::DEVICE);
    forall(ndof, [=] __device__ (int i) -> void
    {
        const double xi[] = {X(0, i), X(1, i)};
        const double c = f_coef(xi);
        d_a2[i] = c * c;
    });
    auto proj = fs.global_indices(MemorySpace::DEVICE);
    forall(fdof, [=] __device__ (int i) -> void
    {
        const int g = proj(i);
        const double xi[] = {X(0, g), X(1, g)};
        d_a[i] = f_coef(xi);
    });

    LinearFunctional l(fem);
    l.action([=] __device__ (const double Xq[2]) -> double {return f_src(Xq, omega);}, d_b);

    Helmholtz A(omega, d_a2, d_a, fem, fs);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::high_resolution_clock::now();
    auto res = gmres(N, d_U, &A, d_b, m, maxit, tol, 0);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::high_resolution_clock::now();

    d.d1("gmres_seconds", 1e-9 * std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
    d.i1("ndof", ndof);
    d.d1("omega", omega);
    d.i1("success", res.success ? 1 : 0);
    d.i1("num_iter", res.num_iter);
    d.i1("num_matvec", res.num_matvec);
    d.dbls("res_norm", res.res_norm.size(), res.res_norm.data());
    d.dbls("b", N, d2h(d_b, N).data());
    d.dbls("U", N, d2h(d_U, N).data());
    return 0;
}

// The examples/DDH.cpp flow at a chosen size.
static int cmd_ddh(int nx, int nb, double omega, int m, int maxit, double tol, uint64_t seed, const std::string & out)
{
    Dump d(out);
    Mesh2D mesh = Mesh2D::uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    const int N = 2 * ndof;

    host_device_dvec U(N), b(N), a(ndof);
    double * d_U = U.device_write();
    double * d_b = b.device_write();
    double * d_a = a.device_write();

    LinearFunctional l(fem);
    DiagInvMassMatrix mi(fem);
    l.action([=] __device__ (const double X[2]) -> double {return f_src(X, omega);}, d_b);
    l.action([] __device__ (const double X[2]) -> double {return f_alpha(X);}, d_a);
    mi.action(d_a, d_a);

    const double * h_a = a.host_read();
    DDH F(omega, h_a, fem, nx, nx);
    const int n_lambda = F.size();

    d.i1("ndof", ndof);
    d.i1("n_lambda", n_lambda);
    d.i1("n_domains", F.n_domains);
    d.i1("nt", F.nt);
    d.d1("dt", F.dt);
    d.i1("mx_dof", F.mx_dof);
    d.i1("mx_fdof", F.mx_fdof);
    d.dbls("a", ndof, h_a);
    d.dbls("b", N, d2h(d_b, N).data());
    d.ints("Bf", F._Bf.size(), F._Bf.host_read());
    d.ints("gI", F._gI.size(), F._gI.host_read());
    d.ints("sI", F._sI.size(), F._sI.host_read());
    d.flts("m", F._m.size(), F._m.host_read());
    d.flts("H", F._H.size(), F._H.host_read());
    d.flts("acoef", F._a.size(), F._a.host_read());
    d.flts("gmi", F._gmi.size(), F._gmi.host_read());
    d.flts("wh_filter", F._wh_filter.size(), F._wh_filter.host_read());
    d.flts("g_tensor", 3 * (int64_t)F._g_tensor.size(), (const float*)F._g_tensor.host_read());

    HostDeviceArray<float> L(n_lambda), Y(n_lambda), T1(n_lambda), T2(n_lambda);
    float * d_L = L.device_write();
    float * d_Y = Y.device_write();
    float * d_T1 = T1.device_write();
    float * d_T2 = T2.device_write();

    F.rhs(d_b, d_Y);
    d.flts("rhs", n_lambda, d2h(d_Y, n_lambda).data());

    {   // one action on a seeded vector (orphan slots hold whatever the reference leaves there;
        // T2 repeats the call so that the reference's own run-to-run spread is recorded)
        std::mt19937 gen((unsigned)seed);
        std::uniform_real_distribution<float> dist(-1.0f, 1.0f);
        std::vector<float> hl(n_lambda);
        for (auto & v : hl) v = dist(gen);
        HostDeviceArray<float> Lr(n_lambda);
        float * d_Lr = Lr.device_write();
        cudaMemcpy(d_Lr, hl.data(), n_lambda * sizeof(float), cudaMemcpyHostToDevice);
        F.action(d_Lr, d_T1);
        F.action(d_Lr, d_T2);
        d.flts("act_x", n_lambda, hl.data());
        d.flts("act_y", n_lambda, d2h(d_T1, n_lambda).data());
        d.flts("act_y2", n_lambda, d2h(d_T2, n_lambda).data());
    }

    auto t0 = std::chrono::high_resolution_clock::now();
    auto res = gmres(n_lambda, d_L, &F, d_Y, m, maxit, (float)tol, 0);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::high_resolution_clock::now();
    F.postprocess(d_L, d_b, d_U);

    d.i1("success", res.success ? 1 : 0);
    d.i1("num_iter", res.num_iter);
    d.i1("num_matvec", res.num_matvec);
    d.d1("gmres_seconds", 1e-9 * std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
    d.dbls("res_norm", res.res_norm.size(), res.res_norm.data());
    d.flts("lambda", n_lambda, d2h(d_L, n_lambda).data());
    d.dbls("U", N, d2h(d_U, N).data());
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    return 0;
}

// timing of the reference's own kernels (CUDA events, default stream)
template <typename F>
static double time_ms(int reps, F && fn)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) fn();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) fn();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / reps;
}

static int cmd_time_ops(int nx, int nb, double omega, int reps)
{
    Mesh2D mesh = Mesh2D::uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    ivec bf = mesh.boundary_edges();
    FaceSpace fs(fem, bf.size(), bf.data());
    const int fdof = fs.size();

    host_device_dvec _x(2 * ndof), _y(2 * ndof), _a2(ndof), _af(fdof);
    double * x = _x.device_write();
    double * y = _y.device_write();
    double * a2 = _a2.device_write();
    double * af = _af.device_write();
    std::vector<double> hx(2 * (size_t)ndof);
    fill_uniform(hx, 12345);
    cudaMemcpy(x, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice);
    fill(ndof, 1.0, a2);
    fill(fdof, 1.0, af);

    double tS, tM, tH, tA;
    {
        StiffnessMatrix S(fem);
        tS = time_ms(reps, [&]() { S.action(x, y); });
    }
    {
        MassMatrix M(a2, fem);
        tM = time_ms(reps, [&]() { M.action(x, y); });
    }
    {
        Helmholtz A(omega, a2, af, fem, fs);
        tA = time_ms(reps, [&]() { A.action(x, y); });
    }
    {
        FaceMassMatrix H(af, fs);
        host_device_dvec _xf(fdof), _yf(fdof);
        double * xf = _xf.device_write();
        double * yf = _yf.device_write();
        tH = time_ms(reps, [&]() { zeros(fdof, yf); fs.restrict(x, xf); H.action(-omega, xf, yf); fs.prolong(yf, y); });
    }
    printf("{\"nx\": %d, \"nb\": %d, \"ndof\": %d, \"stiffness_ms\": %.6f, \"mass_ms\": %.6f, \"face_ms\": %.6f, \"helmholtz_ms\": %.6f}\n",
           nx, nb, ndof, tS, tM, tH, tA);
    return 0;
}

static int cmd_time_ddh(int nx, int nb, double omega, int reps)
{
    Mesh2D mesh = Mesh2D::uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0);
    Basis basis(nb);
    H1Space fem(mesh, basis);
    const int ndof = fem.size();
    std::vector<double> ha(ndof, 1.0);
    DDH F(omega, ha.data(), fem, nx, nx);
    const int n = F.size();
    HostDeviceArray<float> L(n), Y(n);
    float * d_L = L.device_write();
    float * d_Y = Y.device_write();
    fill(n, 0.5f, d_L);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    F.action(d_L, d_Y);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) F.action(d_L, d_Y);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"nx\": %d, \"nb\": %d, \"n_domains\": %d, \"nt\": %d, \"n_lambda\": %d, \"action_ms\": %.6f}\n",
           nx, nb, F.n_domains, F.nt, n, ms / reps);
    return 0;
}

int main(int argc, char ** argv)
{
    if (argc < 2)
    {
        fprintf(stderr,
            "usage: ref_driver tables <nb> <nq> <out>\n"
            "       ref_driver h1 <mesh> <nb> <out>\n"
            "       ref_driver ensemble <nx> <nb> <block> <out>\n"
            "       ref_driver ops <mesh> <nb> <omega> <seed> <out>            (GPU)\n"
            "       ref_driver ops_lite <mesh> <nb> <omega> <seed> <out>       (GPU)\n"
            "       ref_driver metrics_lf <mesh> <nb> <nq> <out>               (GPU)\n"
            "       ref_driver helm_gmres <mesh> <nb> <omega> <m> <maxit> <tol> <out>   (GPU)\n"
            "       ref_driver ddh <nx> <nb> <omega> <m> <maxit> <tol> <seed> <out>     (GPU)\n"
            "       ref_driver time_ops <nx> <nb> <omega> <reps>                (GPU)\n"
            "       ref_driver time_ddh <nx> <nb> <omega> <reps>                (GPU)\n"
            "  <mesh> = rect:<nx> | file:<path>\n");
        return 1;
    }
    const std::string cmd = argv[1];
    if (cmd == "tables" && argc == 5) return cmd_tables(atoi(argv[2]), atoi(argv[3]), argv[4]);
    if (cmd == "h1" && argc == 5) return cmd_h1(argv[2], atoi(argv[3]), argv[4]);
    if (cmd == "ensemble" && argc == 6) return cmd_ensemble(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5]);
    if (cmd == "ops" && argc == 7) return cmd_ops(argv[2], atoi(argv[3]), atof(argv[4]), strtoull(argv[5], nullptr, 10), argv[6]);
    if (cmd == "ops_lite" && argc == 7) return cmd_ops_lite(argv[2], atoi(argv[3]), atof(argv[4]), strtoull(argv[5], nullptr, 10), argv[6]);
    if (cmd == "metrics_lf" && argc == 6) return cmd_metrics_lf(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
    if (cmd == "helm_gmres" && argc == 9) return cmd_helm_gmres(argv[2], atoi(argv[3]), atof(argv[4]), atoi(argv[5]), atoi(argv[6]), atof(argv[7]), argv[8]);
    if (cmd == "ddh" && argc == 10) return cmd_ddh(atoi(argv[2]), atoi(argv[3]), atof(argv[4]), atoi(argv[5]), atoi(argv[6]), atof(argv[7]), strtoull(argv[8], nullptr, 10), argv[9]);
    if (cmd == "time_ops" && argc == 6) return cmd_time_ops(atoi(argv[2]), atoi(argv[3]), atof(argv[4]), atoi(argv[5]));
    if (cmd == "time_ddh" && argc == 6) return cmd_time_ddh(atoi(argv[2]), atoi(argv[3]), atof(argv[4]), atoi(argv[5]));
    fprintf(stderr, "ref_driver: bad command line\n");
    return 1;
}
