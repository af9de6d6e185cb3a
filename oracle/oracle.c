/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the reference's hot-path arithmetic. Each function
 * cites the reference file:line it follows (paths relative to /root/reference). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (cuddhelmholtz_b200/csrc) never does.
 *
 * Parity status: PINNED. scripts/make_golden.py checks this file against the reference's own
 * analytic known-answer tests (tests/mass.cpp, tests/stiffness.cpp) using the reference's own
 * host objects (oracle/_ref) for the index maps and tables, and tests/test_gpu_reference.py
 * compares it with the reference's CUDA kernels run on the B200 box.
 *
 * All arrays are column-major (first index fastest) exactly as include/Tensor.hpp:31-47.
 * Element loops are OpenMP-parallel with a per-thread scatter that is made safe by atomics
 * (`omp atomic`), mirroring the reference's atomicAdd assembly; `orc_set_threads(1)` gives a
 * serial, fixed-order run.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <omp.h>

#define MAXQ 32

void orc_set_threads(int n) { omp_set_num_threads(n); }
int orc_max_threads(void) { return omp_get_max_threads(); }

/* ---- metrics: source/Element.cpp:21-27 (QuadElement::jacobian), include/Element.hpp:52-57 (measure),
 *      source/Element.cpp:5-19 (physical_coordinates); loop shape of source/Mesh2D.cpp:173-197.
 *      corners: (2,4,nel) physical coordinates of the element's CCW corners. */
void orc_element_jacobians(int64_t nel, int nq, const double *xq, const double *corners, double *J /* (2,2,nq,nq,nel) */)
{
#pragma omp parallel for
    for (int64_t el = 0; el < nel; ++el) {
        const double *x = corners + 8 * el; /* x[2*c + d] */
        for (int j = 0; j < nq; ++j)
            for (int i = 0; i < nq; ++i) {
                const double xi0 = xq[i], xi1 = xq[j];
                double *Jp = J + 4 * (i + (int64_t)nq * (j + (int64_t)nq * el));
                Jp[0] = 0.25 * ((1.0 - xi1) * (x[2] - x[0]) + (1.0 + xi1) * (x[4] - x[6]));
                Jp[1] = 0.25 * ((1.0 - xi1) * (x[3] - x[1]) + (1.0 + xi1) * (x[5] - x[7]));
                Jp[2] = 0.25 * ((1.0 - xi0) * (x[6] - x[0]) + (1.0 + xi0) * (x[4] - x[2]));
                Jp[3] = 0.25 * ((1.0 - xi0) * (x[7] - x[1]) + (1.0 + xi0) * (x[5] - x[3]));
            }
    }
}

void orc_element_measures(int64_t nel, int nq, const double *xq, const double *corners, double *detJ /* (nq,nq,nel) */)
{
#pragma omp parallel for
    for (int64_t el = 0; el < nel; ++el) {
        const double *x = corners + 8 * el;
        for (int j = 0; j < nq; ++j)
            for (int i = 0; i < nq; ++i) {
                const double xi0 = xq[i], xi1 = xq[j];
                double Jp[4];
                Jp[0] = 0.25 * ((1.0 - xi1) * (x[2] - x[0]) + (1.0 + xi1) * (x[4] - x[6]));
                Jp[1] = 0.25 * ((1.0 - xi1) * (x[3] - x[1]) + (1.0 + xi1) * (x[5] - x[7]));
                Jp[2] = 0.25 * ((1.0 - xi0) * (x[6] - x[0]) + (1.0 + xi0) * (x[4] - x[2]));
                Jp[3] = 0.25 * ((1.0 - xi0) * (x[7] - x[1]) + (1.0 + xi0) * (x[5] - x[3]));
                detJ[i + (int64_t)nq * (j + (int64_t)nq * el)] = Jp[0] * Jp[3] - Jp[1] * Jp[2];
            }
    }
}

void orc_element_coordinates(int64_t nel, int nq, const double *xq, const double *corners, double *X /* (2,nq,nq,nel) */)
{
#pragma omp parallel for
    for (int64_t el = 0; el < nel; ++el) {
        const double *x = corners + 8 * el;
        for (int j = 0; j < nq; ++j)
            for (int i = 0; i < nq; ++i) {
                const double xi0 = xq[i], xi1 = xq[j];
                const double b[4] = {0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1),
                                     0.25 * (1.0 + xi0) * (1.0 + xi1), 0.25 * (1.0 - xi0) * (1.0 + xi1)};
                double x0 = 0.0, x1 = 0.0;
                for (int c = 0; c < 4; ++c) {
                    x0 += x[2 * c] * b[c];
                    x1 += x[2 * c + 1] * b[c];
                }
                double *Xp = X + 2 * (i + (int64_t)nq * (j + (int64_t)nq * el));
                Xp[0] = x0;
                Xp[1] = x1;
            }
    }
}

/* ---- source/StiffnessMatrix.cpp:5-38 setup_geometric_factors */
void orc_stiffness_setup(int64_t nel, int nq, const double *w, const double *J, double *G /* (3,nq,nq,nel) */)
{
#pragma omp parallel for
    for (int64_t el = 0; el < nel; ++el)
        for (int j = 0; j < nq; ++j)
            for (int i = 0; i < nq; ++i) {
                const double *Jp = J + 4 * (i + (int64_t)nq * (j + (int64_t)nq * el));
                const double W = w[i] * w[j];
                const double Y_eta = Jp[3], X_eta = Jp[2], Y_xi = Jp[1], X_xi = Jp[0];
                const double det = X_xi * Y_eta - X_eta * Y_xi;
                double *Gp = G + 3 * (i + (int64_t)nq * (j + (int64_t)nq * el));
                Gp[0] = W * (Y_eta * Y_eta + X_eta * X_eta) / det;
                Gp[1] = -W * (Y_xi * Y_eta + X_xi * X_eta) / det;
                Gp[2] = W * (Y_xi * Y_xi + X_xi * X_xi) / det;
            }
}

/* ---- source/StiffnessMatrix.cpp:83-184 stiffness_action: out[I] += c * B^T G B u[I].
 * P,D are (nq,nb) column-major: P(q,b) = P[q + nq*b]. Contraction order follows the kernel. */
void orc_stiffness_action(int64_t nel, int nq, int nb, const double *P, const double *D, const double *G,
                          const int *I, double c, const double *u_g, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t el = 0; el < nel; ++el) {
        double u[MAXQ][MAXQ], Pu[MAXQ][MAXQ], Du[MAXQ][MAXQ], F[MAXQ][MAXQ][2];
        const int *Ie = I + (int64_t)nb * nb * el;
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nb; ++tx)
                u[tx][ty] = u_g[Ie[tx + nb * ty]];
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                double pxu = 0.0, dxu = 0.0;
                for (int k = 0; k < nb; ++k) {
                    const double uk = u[k][ty];
                    pxu += P[tx + nq * k] * uk;
                    dxu += D[tx + nq * k] * uk;
                }
                Pu[tx][ty] = pxu;
                Du[tx][ty] = dxu;
            }
        for (int ty = 0; ty < nq; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                const double *Gp = G + 3 * (tx + (int64_t)nq * (ty + (int64_t)nq * el));
                double Dx = 0.0, Dy = 0.0;
                for (int l = 0; l < nb; ++l) {
                    Dx += P[ty + nq * l] * Du[tx][l];
                    Dy += D[ty + nq * l] * Pu[tx][l];
                }
                F[tx][ty][0] = Gp[0] * Dx + Gp[1] * Dy;
                F[tx][ty][1] = Gp[1] * Dx + Gp[2] * Dy;
            }
        for (int ty = 0; ty < nq; ++ty)
            for (int tx = 0; tx < nb; ++tx) {
                double df = 0.0, pg = 0.0;
                for (int i = 0; i < nq; ++i) {
                    df += D[i + nq * tx] * F[i][ty][0];
                    pg += P[i + nq * tx] * F[i][ty][1];
                }
                Du[tx][ty] = df;
                Pu[tx][ty] = pg;
            }
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nb; ++tx) {
                double Su = 0.0;
                for (int j = 0; j < nq; ++j)
                    Su += P[j + nq * ty] * Du[tx][j] + D[j + nq * ty] * Pu[tx][j];
                Su *= c;
#pragma omp atomic
                out[Ie[tx + nb * ty]] += Su;
            }
    }
}

/* ---- source/MassMatrix.cpp:5-67 init_mass_matrix: op = interp(a) * w_i w_j detJ */
void orc_mass_setup(int64_t nel, int nb, int nq, const double *a /* nodal or NULL */, const double *detJ,
                    const double *w, const int *I, const double *P, double *op /* (nq,nq,nel) */)
{
#pragma omp parallel for
    for (int64_t el = 0; el < nel; ++el) {
        double Q[MAXQ][MAXQ], z[MAXQ][MAXQ];
        const int *Ie = I + (int64_t)nb * nb * el;
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nb; ++tx)
                Q[tx][ty] = a ? a[Ie[tx + nb * ty]] : 1.0;
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                double px = 0.0;
                for (int k = 0; k < nb; ++k)
                    px += P[tx + nq * k] * Q[k][ty];
                z[tx][ty] = px;
            }
        for (int ty = 0; ty < nq; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                double ppx = 0.0;
                for (int l = 0; l < nb; ++l)
                    ppx += P[ty + nq * l] * z[tx][l];
                ppx *= w[tx] * w[ty] * detJ[tx + (int64_t)nq * (ty + (int64_t)nq * el)];
                op[tx + (int64_t)nq * (ty + (int64_t)nq * el)] = ppx;
            }
    }
}

/* ---- source/MassMatrix.cpp:137-211 mass_action: y[I] += c * P^T diag(a) P x[I] */
void orc_mass_action(int64_t nel, int nq, int nb, const int *I, const double *P, const double *a,
                     double c, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t el = 0; el < nel; ++el) {
        double u[MAXQ][MAXQ], Pu[MAXQ][MAXQ];
        const int *Ie = I + (int64_t)nb * nb * el;
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nb; ++tx)
                u[tx][ty] = x[Ie[tx + nb * ty]];
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                double pu = 0.0;
                for (int k = 0; k < nb; ++k)
                    pu += P[tx + nq * k] * u[k][ty];
                Pu[tx][ty] = pu;
            }
        for (int ty = 0; ty < nq; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                double ppu = 0.0;
                for (int l = 0; l < nb; ++l)
                    ppu += P[ty + nq * l] * Pu[tx][l];
                u[tx][ty] = a[tx + (int64_t)nq * (ty + (int64_t)nq * el)] * ppu;
            }
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nq; ++tx) {
                double qu = 0.0;
                for (int j = 0; j < nq; ++j)
                    qu += P[j + nq * ty] * u[tx][j];
                Pu[tx][ty] = qu;
            }
        for (int ty = 0; ty < nb; ++ty)
            for (int tx = 0; tx < nb; ++tx) {
                double qqu = 0.0;
                for (int i = 0; i < nq; ++i)
                    qqu += P[i + nq * tx] * Pu[i][ty];
                qqu *= c;
#pragma omp atomic
                y[Ie[tx + nb * ty]] += qqu;
            }
    }
}

/* ---- source/MassMatrix.cpp:241-284 init_diag_mass: op = 1 / sum_e w_i w_j detJ [a] (GLL collocation) */
void orc_diag_inv_mass(int64_t ndof, int64_t nel, int nb, const double *a, const double *detJ /* (nb,nb,nel) at GLL */,
                       const double *w, const int *I, double *op)
{
    memset(op, 0, sizeof(double) * (size_t)ndof);
    for (int64_t el = 0; el < nel; ++el)
        for (int j = 0; j < nb; ++j)
            for (int i = 0; i < nb; ++i) {
                const int idx = I[i + nb * (j + (int64_t)nb * el)];
                double m = w[i] * w[j] * detJ[i + nb * (j + (int64_t)nb * el)];
                if (a)
                    m *= a[idx];
                op[idx] += m;
            }
    for (int64_t i = 0; i < ndof; ++i)
        op[i] = 1.0 / op[i];
}

/* ---- source/FaceMassMatrix.cpp:5-49 init_face_mass */
void orc_facemass_setup(int nf, int nb, int nq, const double *w, const double *P, const double *detJ /* (nq,nf) */,
                        const double *a /* face-space vector or NULL */, const int *I /* (nb,nf) */, double *op /* (nq,nf) */)
{
    for (int e = 0; e < nf; ++e) {
        double al[MAXQ];
        for (int k = 0; k < nb; ++k)
            al[k] = a ? a[I[k + nb * e]] : 1.0;
        for (int k = 0; k < nq; ++k) {
            double pa = 0.0;
            for (int l = 0; l < nb; ++l)
                pa += P[k + nq * l] * al[l];
            pa *= w[k] * detJ[k + nq * e];
            op[k + nq * e] = pa;
        }
    }
}

/* ---- source/FaceMassMatrix.cpp:141-193 mass_action (face) */
void orc_facemass_action(int nf, int nb, int nq, const double *P, const double *a, const int *I, double c,
                         const double *U, double *out)
{
    for (int f = 0; f < nf; ++f) {
        double u[MAXQ], Pu[2 * MAXQ];
        for (int k = 0; k < nb; ++k)
            u[k] = U[I[k + nb * f]];
        for (int k = 0; k < nq; ++k) {
            double pu = 0.0;
            for (int l = 0; l < nb; ++l)
                pu += P[k + nq * l] * u[l];
            pu *= a[k + nq * f];
            Pu[k] = pu;
        }
        for (int k = 0; k < nb; ++k) {
            double Mu = 0.0;
            for (int i = 0; i < nq; ++i)
                Mu += P[i + nq * k] * Pu[i];
            Mu *= c;
            out[I[k + nb * f]] += Mu;
        }
    }
}

/* ---- source/FaceMassMatrix.cpp:226-270 init_diag: lumped face mass at GLL nodes, reciprocal */
void orc_diag_inv_facemass(int fdof, int nf, int nb, const double *w, const double *detJ /* (nb,nf) */, const double *a,
                           const int *I, double *op)
{
    memset(op, 0, sizeof(double) * (size_t)fdof);
    for (int f = 0; f < nf; ++f)
        for (int i = 0; i < nb; ++i) {
            const int idx = I[i + nb * f];
            double m = w[i] * detJ[i + nb * f];
            if (a)
                m *= a[idx];
            op[idx] += m;
        }
    for (int i = 0; i < fdof; ++i)
        op[i] = 1.0 / op[i];
}

/* ---- source/H1Space.cpp:189-219 FaceSpace::restrict / prolong / orth */
void orc_restrict(int fdof, const int *proj, const double *x, double *y)
{
    for (int i = 0; i < fdof; ++i)
        y[i] = x[proj[i]];
}
void orc_prolong(int fdof, const int *proj, const double *x, double *y)
{
    for (int i = 0; i < fdof; ++i)
        y[proj[i]] += x[i];
}
void orc_orth(int fdof, const int *proj, double *x)
{
    for (int i = 0; i < fdof; ++i)
        x[proj[i]] = 0.0;
}

/* ---- examples/Helmholtz.hpp:28-56 composite:
 *   Au = S u - w^2 M u - w H v ;  Av = -(S v - w^2 M v + w H u)        (x = [u; v], y = [Au; Av]) */
void orc_helmholtz_action(int64_t ndof, int fdof, int64_t nel, int nf, int nb, int nqS, int nqM, int nqH,
                          const int *I, const double *PS, const double *DS, const double *G,
                          const double *PM, const double *aM,
                          const int *If, const int *proj, const double *PH, const double *aH,
                          double omega, const double *x, double *y, double *xf, double *yf)
{
    const double *u = x, *v = x + ndof;
    double *Au = y, *Av = y + ndof;
    memset(y, 0, sizeof(double) * 2 * (size_t)ndof);
    orc_stiffness_action(nel, nqS, nb, PS, DS, G, I, 1.0, u, Au);
    orc_stiffness_action(nel, nqS, nb, PS, DS, G, I, 1.0, v, Av);
    orc_mass_action(nel, nqM, nb, I, PM, aM, -omega * omega, u, Au);
    orc_mass_action(nel, nqM, nb, I, PM, aM, -omega * omega, v, Av);

    memset(yf, 0, sizeof(double) * (size_t)fdof);
    orc_restrict(fdof, proj, v, xf);
    orc_facemass_action(nf, nb, nqH, PH, aH, If, -omega, xf, yf);
    orc_prolong(fdof, proj, yf, Au);

    memset(yf, 0, sizeof(double) * (size_t)fdof);
    orc_restrict(fdof, proj, u, xf);
    orc_facemass_action(nf, nb, nqH, PH, aH, If, omega, xf, yf);
    orc_prolong(fdof, proj, yf, Av);

    for (int64_t i = 0; i < ndof; ++i)
        Av[i] *= -1.0;
}

/* ---- source/DDH.cpp:31-58 init_geom_factors: FP64 -> float3 at GLL nodes, per subdomain element slot */
void orc_ddh_geom(int n_domains, int mx_elems, int nb, const int *n_elems, const int *elems /* (mx_elems,dom) */,
                  const double *w, const double *J /* (2,2,nb,nb,g_elem) */, float *G /* (3,nb,nb,mx_elems,dom) */)
{
    memset(G, 0, sizeof(float) * 3 * (size_t)nb * nb * mx_elems * n_domains);
    for (int s = 0; s < n_domains; ++s)
        for (int el = 0; el < n_elems[s]; ++el) {
            const int64_t g_el = elems[el + (int64_t)mx_elems * s];
            for (int j = 0; j < nb; ++j)
                for (int i = 0; i < nb; ++i) {
                    const double *Jp = J + 4 * (i + (int64_t)nb * (j + (int64_t)nb * g_el));
                    const double W = w[i] * w[j];
                    const double Y_eta = Jp[3], X_eta = Jp[2], Y_xi = Jp[1], X_xi = Jp[0];
                    const double det = X_xi * Y_eta - X_eta * Y_xi;
                    float *g = G + 3 * (i + (int64_t)nb * (j + (int64_t)nb * (el + (int64_t)mx_elems * s)));
                    g[0] = (float)(W * (Y_eta * Y_eta + X_eta * X_eta) / det);
                    g[1] = (float)(-W * (Y_xi * Y_eta + X_xi * X_eta) / det);
                    g[2] = (float)(W * (Y_xi * Y_xi + X_xi * X_xi) / det);
                }
        }
}

/* ---- source/DDH.cpp:60-109 stiffness<NB> (device function), emulated for all MX threads of one CTA.
 * s_I: [el][l][k] -> subspace DOF; s_D: [k][l] = D(k,l); s_u input (indexed by DOF) and work; s_w work;
 * s_out accumulates by DOF. Threads are visited in tid order (one valid atomic ordering). */
static void ddh_stiffness(int nb, int mx, const float *g3, const int *s_I, const float *s_D, float *s_w, float *s_u,
                          float *s_out, float *Ux, float *Uy)
{
    const int nb2 = nb * nb;
    for (int tid = 0; tid < mx; ++tid) {
        const int k = tid % nb, l = (tid % nb2) / nb, el = tid / nb2;
        float ux = 0.0f, uy = 0.0f;
        for (int i = 0; i < nb; ++i) {
            int idx = s_I[i + nb * (l + nb * el)];
            ux += s_D[k * nb + i] * s_u[idx];
            idx = s_I[k + nb * (i + nb * el)];
            uy += s_D[l * nb + i] * s_u[idx];
        }
        Ux[tid] = ux;
        Uy[tid] = uy;
    }
    for (int tid = 0; tid < mx; ++tid) {
        const float *G = g3 + 3 * tid;
        s_u[tid] = G[0] * Ux[tid] + G[1] * Uy[tid];
        s_w[tid] = G[1] * Ux[tid] + G[2] * Uy[tid];
    }
    for (int tid = 0; tid < mx; ++tid) {
        const int k = tid % nb, l = (tid % nb2) / nb, el = tid / nb2;
        float Su = 0.0f;
        for (int i = 0; i < nb; ++i) {
            int idx = i + nb * (l + nb * el);
            Su += s_D[i * nb + k] * s_u[idx];
            idx = k + nb * (i + nb * el);
            Su += s_D[i * nb + l] * s_w[idx];
        }
        s_out[s_I[k + nb * (l + nb * el)]] += Su;
    }
}

/* ---- source/DDH.cpp:111-321 ddh_action<NB,NEL> (+ :611-695 the three entry points choose x / y / lambda / update).
 * mx = NB*NB*NEL*NEL threads per subdomain (= block*block). Arrays as built by DDH::DDH (:323-609):
 *   B (mx_fdof,2,dom), gI (mx_dof,dom), sI (nb,nb,mx_elems,dom), D (nb,nb) float, g (3, mx, dom) float,
 *   m, gmi, a (mx_dof,dom), H (mx_fdof,dom), wh_filter[nt+1], cs/sn[2nt+1].
 * y (if non-NULL) must be zeroed by the caller (the reference zeroes it at :146). */
void orc_ddh_action(int n_domains, int nb, int mx, int mx_dof, int mx_fdof, int64_t g_ndof, int64_t n_lambda, int nt,
                    float omega, float dt, const int *s_dof, const int *s_fdof, const int *B, const int *gI, const int *sI,
                    const float *D, const float *g, const float *m, const float *gmi, const float *a, const float *H,
                    const float *wh_filter, const float *cs, const float *sn, const double *x, double *y,
                    const float *d_lambda, float *d_update)
{
    const float half_dt = 0.5f * dt;
    const float rw = 1.0f / omega;
    const float *g_lambda = d_lambda, *g_mu = d_lambda ? d_lambda + n_lambda : NULL;
    float *lambda_update = d_update, *mu_update = d_update ? d_update + n_lambda : NULL;
    const int mx_el = mx / (nb * nb);

#pragma omp parallel for schedule(dynamic, 1)
    for (int subsp = 0; subsp < n_domains; ++subsp) {
        const int fdof = s_fdof[subsp], ndof = s_dof[subsp];
        float *buf = (float *)calloc((size_t)mx * 18, sizeof(float));
        float *s_p_half = buf, *s_q_half = buf + mx, *s_z = buf + 2 * mx, *Ux = buf + 3 * mx, *Uy = buf + 4 * mx;
        float *ai = buf + 5 * mx, *mi = buf + 6 * mx, *inv_mi = buf + 7 * mx, *Hi = buf + 8 * mx, *F = buf + 9 * mx,
              *G = buf + 10 * mx, *u = buf + 11 * mx, *v = buf + 12 * mx, *p = buf + 13 * mx, *q = buf + 14 * mx,
              *lam = buf + 15 * mx, *mu = buf + 16 * mx;
        int *s_I = (int *)malloc(sizeof(int) * (size_t)mx);
        float s_D[MAXQ * MAXQ];
        for (int l = 0; l < nb; ++l)
            for (int k = 0; k < nb; ++k)
                s_D[k * nb + l] = D[k + nb * l];
        for (int tid = 0; tid < mx; ++tid)
            s_I[tid] = sI[tid + (int64_t)nb * nb * mx_el * subsp];
        const float *g3 = g + 3 * (int64_t)mx * subsp;

        for (int tid = 0; tid < ndof; ++tid) {
            ai[tid] = a[tid + (int64_t)mx_dof * subsp];
            mi[tid] = m[tid + (int64_t)mx_dof * subsp];
            inv_mi[tid] = 1.0f / (ai[tid] * ai[tid] * mi[tid]);
            if (x) {
                const int64_t gi = gI[tid + (int64_t)mx_dof * subsp];
                F[tid] = (float)x[gi];
                G[tid] = (float)x[g_ndof + gi];
            }
        }
        for (int tid = 0; tid < fdof; ++tid) {
            Hi[tid] = H[tid + (int64_t)mx_fdof * subsp];
            if (d_lambda) {
                const int idx = B[tid + mx_fdof * (0 + 2 * (int64_t)subsp)];
                if (idx >= 0) {
                    lam[tid] = g_lambda[idx];
                    mu[tid] = g_mu[idx];
                    F[tid] += Hi[tid] * lam[tid];
                    G[tid] += Hi[tid] * mu[tid];
                }
            }
            Hi[tid] *= ai[tid];
        }

        for (int whit = 0; whit < 5; ++whit) {
            float dK = wh_filter[0];
            for (int tid = 0; tid < mx; ++tid) {
                p[tid] = u[tid];
                q[tid] = v[tid];
                u[tid] *= dK;
                v[tid] *= dK;
            }
            for (int it = 1; it <= nt; ++it) {
                for (int tid = 0; tid < mx; ++tid) {
                    s_z[tid] = 0.0f;
                    s_p_half[tid] = p[tid];
                }
                ddh_stiffness(nb, mx, g3, s_I, s_D, s_q_half, s_p_half, s_z, Ux, Uy);
                for (int tid = 0; tid < mx; ++tid) {
                    s_z[tid] -= Hi[tid] * q[tid];
                    float dq = s_z[tid] + cs[2 * it - 2] * F[tid];
                    dq += sn[2 * it - 2] * G[tid];
                    dq *= inv_mi[tid];
                    s_p_half[tid] = p[tid] - half_dt * q[tid];
                    s_q_half[tid] = q[tid] + half_dt * dq;
                    s_z[tid] = 0.0f;
                    p[tid] -= dt * s_q_half[tid];
                    s_z[tid] -= Hi[tid] * s_q_half[tid];
                }
                ddh_stiffness(nb, mx, g3, s_I, s_D, s_q_half, s_p_half, s_z, Ux, Uy);
                dK = wh_filter[it];
                for (int tid = 0; tid < mx; ++tid) {
                    float dq = s_z[tid] + cs[2 * it - 1] * F[tid];
                    dq += sn[2 * it - 1] * G[tid];
                    dq *= inv_mi[tid];
                    q[tid] += dt * dq;
                    u[tid] += dK * p[tid];
                    v[tid] += dK * q[tid];
                }
            }
        }

        for (int tid = 0; tid < mx; ++tid)
            v[tid] *= rw;

        if (y)
            for (int tid = 0; tid < ndof; ++tid) {
                const int64_t gi = gI[tid + (int64_t)mx_dof * subsp];
                const float M = mi[tid] * gmi[tid + (int64_t)mx_dof * subsp];
                const double m_u = M * u[tid];
                const double m_v = M * v[tid];
#pragma omp atomic
                y[gi] += m_u;
#pragma omp atomic
                y[g_ndof + gi] += m_v;
            }
        if (d_update)
            for (int tid = 0; tid < fdof; ++tid) {
                const int idx = B[tid + mx_fdof * (1 + 2 * (int64_t)subsp)];
                if (idx >= 0) {
                    const float S = 2.0f * ai[tid] * omega;
                    lambda_update[idx] = -lam[tid] - S * v[tid];
                    mu_update[idx] = -mu[tid] + S * u[tid];
                }
            }
        free(buf);
        free(s_I);
    }
}
