// H1 space (global-to-local map + nodal coordinates), face space, and the CTA assembly plan.
//
// Numbering semantics = reference source/H1Space.cpp:11-127: a node shared between elements belongs to
// the lowest-numbered element touching it (edge->elements[0] / connected_elements[0]); global ids are
// handed out in increasing volume index v = i + nb*(j + nb*el) over the owned nodes. That is exactly a
// first-touch scan, which is how it is computed here (flat per-vertex / per-edge id tables, no hash maps).
// Nodal coordinates: bilinear map of the GLL tensor nodes (source/Element.cpp:5-19), last writer wins
// in element order, as in the reference loop (:108-126).
//
// Compile without FP contraction (-ffp-contract=off) so xy is bit-identical to the reference's host pass.
#include "common.hpp"
#include <atomic>
#include <cmath>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

namespace cb200
{
    H1Space::H1Space(const Mesh * mesh_, int nb_) : mesh(mesh_), basis(new Basis(nb_)), nb(nb_), n_elem(mesh_->n_elem)
    {
        CB_REQUIRE(nb >= 2, "H1Space: n_basis must be >= 2");
        const int64_t nel = n_elem;
        const int ne_int = nb - 2;
        CB_REQUIRE((double)nel * nb * nb < 2.0e9, "H1Space: n_elem * n_basis^2 exceeds 32-bit DOF ids");
        I.assign((size_t)nb * nb * nel, -1);
        const double * q = basis->x.data();
        const bool closed_form = mesh->nx > 0 && !(getenv("CUDDH_B200_CLOSED_FORM") && atoi(getenv("CUDDH_B200_CLOSED_FORM")) == 0);
        if (closed_form) {
            // Mesh2D::uniform_rect: the first-touch scan below has a closed form (oracle/setup_np.py:uniform_rect_closed_form, pinned to
            // the reference's arrays up to 1024^2): element (ex, ey) introduces the nodes with (i > 0 or ex == 0) and (j > 0 or ey == 0),
            // numbered in (j, i) order after all nodes of the elements before it; every other node belongs to the element to its left /
            // below. Nodal coordinates: the LAST element that touches a node wins in the reference's loop, i.e. an element writes node
            // (i, j) unless a right / upper neighbour also holds it. Both loops run on all host threads (disjoint writes).
            const int64_t nx = mesh->nx, ny = mesh->ny;
            const int p = nb - 1;
            std::vector<int64_t> base((size_t)nel + 1, 0); // first global id introduced by element el
            for (int64_t el = 0; el < nel; ++el) {
                const int64_t ex = el % nx, ey = el / nx;
                base[el + 1] = base[el] + (int64_t)(ex == 0 ? nb : p) * (ey == 0 ? nb : p);
            }
            ndof = base[nel];
            auto gid = [&](int64_t gx, int64_t gy) -> int {
                const int64_t ox = gx == 0 ? 0 : (gx - 1) / p, oy = gy == 0 ? 0 : (gy - 1) / p; // owning element
                const int64_t oi = gx - p * ox, oj = gy - p * oy;                              // its local node
                const int64_t ni = ox == 0 ? nb : p;
                return (int)(base[ox + nx * oy] + (oj - (oy > 0 ? 1 : 0)) * ni + (oi - (ox > 0 ? 1 : 0)));
            };
            xy.assign(2 * (size_t)ndof, 0.0);
            parallel_for(ny, [&](int64_t jb, int64_t je, int) {
                for (int64_t ey = jb; ey < je; ++ey)
                    for (int64_t ex = 0; ex < nx; ++ex) {
                        const int64_t el = ex + nx * ey;
                        int * Ie = &I[(size_t)nb * nb * el];
                        double c[8];
                        mesh->corners(el, c);
                        for (int j = 0; j < nb; ++j)
                            for (int i = 0; i < nb; ++i) {
                                const int id = gid(ex * p + i, ey * p + j);
                                Ie[i + nb * j] = id;
                                if ((i < p || ex == nx - 1) && (j < p || ey == ny - 1)) { // this element is the last one to touch the node
                                    const double xi0 = q[i], xi1 = q[j];
                                    const double b[4] = {0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1),
                                                         0.25 * (1.0 + xi0) * (1.0 + xi1), 0.25 * (1.0 - xi0) * (1.0 + xi1)};
                                    double x0 = 0.0, x1 = 0.0;
                                    for (int k = 0; k < 4; ++k) {
                                        x0 += c[2 * k] * b[k];
                                        x1 += c[2 * k + 1] * b[k];
                                    }
                                    xy[2 * (size_t)id] = x0;
                                    xy[2 * (size_t)id + 1] = x1;
                                }
                            }
                    }
            });
            return;
        }
        std::vector<int> vertex_id((size_t)mesh->n_nodes, -1);
        std::vector<int> edge_id((size_t)mesh->n_edges * (size_t)std::max(ne_int, 0), -1);
        const int * elems = mesh->elems.data();
        const int * edges = mesh->edges.data();
        const int * eledge = mesh->elem_edges.data();

        int next = 0;
        for (int64_t el = 0; el < nel; ++el) {
            int * Ie = &I[(size_t)nb * nb * el];
            for (int j = 0; j < nb; ++j) {
                const bool jlo = (j == 0), jhi = (j == nb - 1);
                for (int i = 0; i < nb; ++i) {
                    const bool ilo = (i == 0), ihi = (i == nb - 1);
                    int id;
                    if ((ilo || ihi) && (jlo || jhi)) { // corner -> mesh vertex
                        const int c = jlo ? (ilo ? 0 : 1) : (ihi ? 2 : 3);
                        int & vid = vertex_id[elems[4 * el + c]];
                        if (vid < 0)
                            vid = next++;
                        id = vid;
                    }
                    else if (ilo || ihi || jlo || jhi) { // edge-interior node
                        const int s = jlo ? 0 : (ihi ? 1 : (jhi ? 2 : 3));
                        const int pos = (s == 0 || s == 2) ? i : j;
                        const int e = eledge[4 * el + s];
                        const int * rec = edges + 8 * (size_t)e;
                        const bool first = (rec[2] == (int)el && rec[4] == s);
                        const int t = (first || rec[6] > 0) ? pos : (nb - 1 - pos);
                        int & eid = edge_id[(size_t)e * ne_int + (t - 1)];
                        if (eid < 0)
                            eid = next++;
                        id = eid;
                    }
                    else
                        id = next++;
                    Ie[i + nb * j] = id;
                }
            }
        }
        ndof = next;

        xy.assign(2 * (size_t)ndof, 0.0);
        for (int64_t el = 0; el < nel; ++el) {
            double c[8];
            mesh->corners(el, c);
            const int * Ie = &I[(size_t)nb * nb * el];
            for (int j = 0; j < nb; ++j)
                for (int i = 0; i < nb; ++i) {
                    const double xi0 = q[i], xi1 = q[j];
                    const double b[4] = {0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1),
                                         0.25 * (1.0 + xi0) * (1.0 + xi1), 0.25 * (1.0 - xi0) * (1.0 + xi1)};
                    double x0 = 0.0, x1 = 0.0;
                    for (int k = 0; k < 4; ++k) {
                        x0 += c[2 * k] * b[k];
                        x1 += c[2 * k + 1] * b[k];
                    }
                    const size_t idx = (size_t)Ie[i + nb * j];
                    xy[2 * idx] = x0;
                    xy[2 * idx + 1] = x1;
                }
        }
    }

    const int * H1Space::device_I()
    {
        if (!d_I.p)
            d_I.upload(I);
        return d_I.p;
    }
    const double * H1Space::device_xy()
    {
        if (!d_xy.p)
            d_xy.upload(xy);
        return d_xy.p;
    }
    bool H1Space::all_affine()
    {
        if (affine_state < 0) {
            // x0 - x1 + x2 - x3 == 0 (both coordinates) up to rounding of the vertex coordinates
            std::atomic<int> bad{0};
            const Mesh & m = *mesh;
            parallel_for(m.n_elem, [&](int64_t b, int64_t e, int) {
                for (int64_t el = b; el < e && !bad.load(std::memory_order_relaxed); ++el) {
                    double c[8];
                    m.corners(el, c);
                    const double ex = c[0] - c[2] + c[4] - c[6], ey = c[1] - c[3] + c[5] - c[7];
                    const double h = std::fabs(c[2] - c[0]) + std::fabs(c[3] - c[1]) + std::fabs(c[6] - c[0]) + std::fabs(c[7] - c[1]);
                    if (std::fabs(ex) + std::fabs(ey) > 1e-14 * h)
                        bad.store(1, std::memory_order_relaxed);
                }
            });
            affine_state = bad.load() ? 0 : 1;
        }
        return affine_state == 1;
    }

    const double * H1Space::device_corners()
    {
        if (!d_corners.p) {
            std::vector<double> c(8 * (size_t)n_elem);
            for (int64_t el = 0; el < n_elem; ++el)
                mesh->corners(el, &c[8 * (size_t)el]);
            d_corners.upload(c);
        }
        return d_corners.p;
    }
    Plan & H1Space::get_plan_host()
    {
        if (!plan) {
            plan.reset(new Plan);
            build_plan(*this, *plan);
        }
        return *plan;
    }
    Plan & H1Space::get_plan()
    {
        if (!plan) {
            plan.reset(new Plan);
            build_plan(*this, *plan);
        }
        plan->ensure_device();
        return *plan;
    }

    Plan & H1Space::get_plan_tpe()
    {
        if (!plan_tpe) {
            plan_tpe.reset(new Plan);
            build_plan(*this, *plan_tpe, true);
        }
        plan_tpe->ensure_device();
        return *plan_tpe;
    }

    // ---- FaceSpace: reference source/H1Space.cpp:129-187 ----
    FaceSpace::FaceSpace(H1Space * fem_, int64_t nf, const int * faces_) : fem(fem_), nb(fem_->nb), n_faces(nf)
    {
        const Mesh & mesh = *fem->mesh;
        faces.assign(faces_, faces_ + nf);
        I.assign((size_t)nb * nf, -1);
        std::vector<int> seen((size_t)fem->ndof, -1);
        std::vector<double> meas(nf);
        for (int64_t f = 0; f < nf; ++f) {
            const int e = faces[f];
            CB_REQUIRE(e >= 0 && e < mesh.n_edges, "FaceSpace: face index out of range");
            const int * rec = &mesh.edges[8 * (size_t)e];
            const int64_t el = rec[2];
            const int s = rec[4];
            meas[f] = mesh.edge_meas[e];
            for (int i = 0; i < nb; ++i) {
                const int m = (s == 0 || s == 2) ? i : (s == 1 ? nb - 1 : 0);
                const int n = (s == 1 || s == 3) ? i : (s == 2 ? nb - 1 : 0);
                const int g = fem->I[(size_t)m + (size_t)nb * (n + (size_t)nb * el)];
                if (seen[g] < 0) {
                    seen[g] = (int)proj.size();
                    proj.push_back(g);
                }
                I[i + (size_t)nb * f] = seen[g];
            }
        }
        fdof = (int64_t)proj.size();

        // DOF-centric incidence lists in (face, k) order: fixed summation order for the face mass action
        inc_ptr.assign((size_t)fdof + 1, 0);
        for (size_t k = 0; k < I.size(); ++k)
            inc_ptr[I[k] + 1]++;
        for (int64_t d = 0; d < fdof; ++d)
            inc_ptr[d + 1] += inc_ptr[d];
        inc.resize(I.size());
        std::vector<int> cur(inc_ptr.begin(), inc_ptr.end() - 1);
        for (size_t k = 0; k < I.size(); ++k)
            inc[cur[I[k]]++] = (int)k;

        h_meas = meas;
    }

    void FaceSpace::ensure_device()
    {
        if (on_device)
            return;
        d_I.upload(I);
        d_proj.upload(proj);
        d_inc_ptr.upload(inc_ptr);
        d_inc.upload(inc);
        d_meas.upload(h_meas);
        on_device = true;
    }

    // ---- assembly plan -------------------------------------------------------------------------
    // Elements are grouped into CTA patches (2-D tiles on uniform_rect meshes, chunks of a Morton-ordered
    // centroid sort otherwise). Inside a patch the kernel accumulates element results into a patch-local
    // vector in shared memory, colour by colour (elements of one colour share no DOF), so the summation
    // order of every DOF is fixed by the plan -> bitwise reproducible without atomics. DOFs touched by a
    // single patch are written straight to y; DOFs on patch boundaries go to a partial buffer whose slots
    // are summed in patch order by a second small kernel.
    namespace
    {
        void pick_patch_shape(int nb, int & px, int & py)
        {
            switch (nb) {
            case 2: px = 12; py = 10; break;
            case 3: px = 12; py = 10; break;
            case 4: px = 10; py = 6; break;
            case 5: px = 10; py = 6; break;
            case 6: px = 6; py = 5; break;
            case 7: px = 6; py = 5; break;
            case 8: px = 4; py = 3; break;
            case 9: px = 4; py = 3; break;
            default: px = 3; py = 2; break;
            }
        }

        uint64_t morton(uint32_t a, uint32_t b)
        {
            auto spread = [](uint64_t v) {
                v &= 0xffffffffull;
                v = (v | (v << 16)) & 0x0000ffff0000ffffull;
                v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
                v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
                v = (v | (v << 2)) & 0x3333333333333333ull;
                v = (v | (v << 1)) & 0x5555555555555555ull;
                return v;
            };
            return spread(a) | (spread(b) << 1);
        }
    } // namespace

    void build_plan(H1Space & fem, Plan & plan, bool tpe)
    {
        const Mesh & mesh = *fem.mesh;
        const int nb = fem.nb, nb2 = nb * nb;
        const int64_t nel = fem.n_elem;
        int px, py;
        pick_patch_shape(nb, px, py);
        if (tpe) { // one thread per element: a patch is one warpgroup (128 threads) of elements, 8 x 16 tiles (measured: 16 x 8,
                   // 32 x 4, 4 x 32 and 64 x 2 are within 1-5 % of it)
            px = 8;
            py = nb >= 6 ? 8 : 16; // n_basis >= 6: thread-PAIR-per-element kernel (volume_action_pair), 64 elements per patch
            if (const char * e = getenv("CUDDH_B200_TPE_PX"))
                px = std::max(1, atoi(e));
            if (const char * e = getenv("CUDDH_B200_TPE_PY"))
                py = std::max(1, atoi(e));
            CB_REQUIRE((px * py) % 32 == 0, "thread-per-element plan: patch size must be a multiple of 32");
        }
        else {
            if (const char * e = getenv("CUDDH_B200_PX")) // experiment knobs (scripts/sweep_ops.py)
                px = std::max(1, atoi(e));
            if (const char * e = getenv("CUDDH_B200_PY"))
                py = std::max(1, atoi(e));
        }
        const int PE = px * py;
        plan.nb = nb;
        plan.PE = PE;
        plan.node_major = tpe;

        // 1. patches -> element lists
        std::vector<std::vector<int>> pel;
        if (mesh.nx > 0 && (int64_t)mesh.nx * mesh.ny == nel) {
            const int tx = (mesh.nx + px - 1) / px, ty = (mesh.ny + py - 1) / py;
            pel.resize((size_t)tx * ty);
            for (int j = 0; j < mesh.ny; ++j)
                for (int i = 0; i < mesh.nx; ++i)
                    pel[(size_t)(i / px) + (size_t)tx * (j / py)].push_back(i + mesh.nx * j);
        }
        else {
            double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
            for (int64_t v = 0; v < mesh.n_nodes; ++v)
                for (int d = 0; d < 2; ++d) {
                    lo[d] = std::min(lo[d], mesh.xy[2 * v + d]);
                    hi[d] = std::max(hi[d], mesh.xy[2 * v + d]);
                }
            std::vector<std::pair<uint64_t, int>> key(nel);
            for (int64_t el = 0; el < nel; ++el) {
                double c[8];
                mesh.corners(el, c);
                const double cx = 0.25 * (c[0] + c[2] + c[4] + c[6]), cy = 0.25 * (c[1] + c[3] + c[5] + c[7]);
                const uint32_t a = (uint32_t)(65535.0 * (cx - lo[0]) / std::max(hi[0] - lo[0], 1e-300));
                const uint32_t b = (uint32_t)(65535.0 * (cy - lo[1]) / std::max(hi[1] - lo[1], 1e-300));
                key[el] = {morton(a, b), (int)el};
            }
            std::sort(key.begin(), key.end());
            pel.resize((size_t)((nel + PE - 1) / PE));
            for (int64_t k = 0; k < nel; ++k)
                pel[k / PE].push_back(key[k].second);
        }
        const int64_t np = (int64_t)pel.size();
        plan.n_patches = np;
        if (tpe && nb >= 4) { // element-interior ids an affine function of (i, j)? (first-touch numbering: always; checked, not assumed)
            bool ok = true;
            for (int64_t el = 0; el < nel && ok; ++el) {
                const int * Ie = &fem.I[(size_t)nb2 * el];
                const int base = Ie[1 + nb], stride = Ie[1 + 2 * nb] - base;
                for (int j = 1; j < nb - 1 && ok; ++j)
                    for (int i = 1; i < nb - 1; ++i)
                        if (Ie[i + nb * j] != base + (i - 1) + (j - 1) * stride) {
                            ok = false;
                            break;
                        }
            }
            plan.interior_affine = ok;
        }

        // 2. element slots of each patch (patch order = element order of the lists above)
        plan.hdr.resize(np);
        plan.slot_elem.assign((size_t)np * PE, -1);
        for (int64_t p = 0; p < np; ++p) {
            PatchHdr & h = plan.hdr[p];
            h.elem_begin = (int)(p * PE);
            h.n_elem = (int)pel[p].size();
            h.reserved = 0;
            for (size_t k = 0; k < pel[p].size(); ++k)
                plan.slot_elem[(size_t)p * PE + k] = pel[p][k];
        }

        // 3. patch-local DOF lists; a DOF touched by more than one patch is "shared"
        // node (i, j) strictly inside an element belongs to that element alone on any conforming mesh
        auto interior = [nb](int a) { const int i = a % nb, j = a / nb; return i > 0 && i < nb - 1 && j > 0 && j < nb - 1; };
        const int NG = (nb2 + 3) / 4;
        std::vector<uint8_t> touch((size_t)fem.ndof, 0);
        std::vector<std::vector<int>> pg(np);
        for (int64_t p = 0; p < np; ++p) {
            auto & g = pg[p];
            g.reserve((size_t)plan.hdr[p].n_elem * nb2);
            for (int k = 0; k < plan.hdr[p].n_elem; ++k) {
                const int * Ie = &fem.I[(size_t)nb2 * plan.slot_elem[(size_t)p * PE + k]];
                if (plan.node_major) { // element-interior nodes are written by the compute threads themselves
                    for (int a = 0; a < nb2; ++a)
                        if (!interior(a))
                            g.push_back(Ie[a]);
                }
                else
                    g.insert(g.end(), Ie, Ie + nb2);
            }
            std::sort(g.begin(), g.end());
            g.erase(std::unique(g.begin(), g.end()), g.end());
            for (int v : g)
                if (touch[v] < 255)
                    touch[v]++;
        }
        // shared DOF table (ascending global id) and CSR of partial slots in patch order
        std::vector<int> sh_index((size_t)fem.ndof, -1);
        plan.sh_gid.clear();
        for (int64_t g = 0; g < fem.ndof; ++g)
            if (touch[g] > 1) {
                sh_index[g] = (int)plan.sh_gid.size();
                plan.sh_gid.push_back((int)g);
            }
        plan.n_shared = (int64_t)plan.sh_gid.size();
        plan.sh_ptr.assign((size_t)plan.n_shared + 1, 0);
        for (int64_t p = 0; p < np; ++p)
            for (int v : pg[p])
                if (sh_index[v] >= 0)
                    plan.sh_ptr[sh_index[v] + 1]++;
        for (int64_t s = 0; s < plan.n_shared; ++s)
            plan.sh_ptr[s + 1] += plan.sh_ptr[s];
        plan.n_slots_total = plan.sh_ptr[plan.n_shared];
        std::vector<int> next_slot(plan.sh_ptr.begin(), plan.sh_ptr.end() - 1);

        plan.gid.clear();
        plan.slot.clear();
        plan.L.assign((size_t)np * PE * nb2, 0);
        plan.cent.assign((size_t)np * PE * nb2, 0);
        if (plan.node_major) {
            plan.Ig.assign((size_t)np * PE * NG * 4, 0);
        }
        plan.cptr.clear();
        plan.cptr.reserve(plan.gid.capacity() + (size_t)np);
        std::vector<int> local((size_t)fem.ndof, -1);
        std::vector<int> cnt;
        for (int64_t p = 0; p < np; ++p) {
            PatchHdr & h = plan.hdr[p];
            auto & g = pg[p];
            h.pdof_begin = (int)plan.gid.size();
            h.slot_begin = (int)plan.slot.size();
            h.n_pdof = (int)g.size();
            CB_REQUIRE(g.size() < 65534, "assembly plan: patch has too many DOFs for 16-bit local ids");
            int n = 0;
            for (int v : g)
                if (sh_index[v] < 0) {
                    local[v] = n++;
                    plan.gid.push_back(v);
                }
            h.n_int = n;
            for (int v : g)
                if (sh_index[v] >= 0) {
                    local[v] = n++;
                    plan.gid.push_back(v);
                    plan.slot.push_back(next_slot[sh_index[v]]++);
                }
            plan.max_pdof = std::max(plan.max_pdof, h.n_pdof);
            plan.max_nsh = std::max(plan.max_nsh, h.n_pdof - h.n_int);
            // entry id of (slot k, node a): element-major k*nb2 + a, or node-major a*PE + k (thread-per-element kernels)
            auto entry = [&](int k, int a) { return plan.node_major ? a * PE + k : k * nb2 + a; };
            for (int k = 0; k < h.n_elem; ++k) {
                const int * Ie = &fem.I[(size_t)nb2 * plan.slot_elem[(size_t)p * PE + k]];
                uint16_t * Lp = &plan.L[(size_t)p * PE * nb2];
                for (int a = 0; a < nb2; ++a)
                    Lp[entry(k, a)] = (plan.node_major && interior(a)) ? (uint16_t)0xFFFF : (uint16_t)local[Ie[a]];
                if (plan.node_major) {
                    for (int a = 0; a < nb2; ++a) // groups of four nodes per 16-byte load: ((p, a / 4), slot, a % 4)
                        plan.Ig[(((size_t)p * NG + a / 4) * PE + k) * 4 + a % 4] = Ie[a];
                }
            }
            // CSR: patch-local DOF -> its element-local entries in ascending (slot, node) order. This fixes the
            // summation order of every DOF (deterministic assembly without atomics or colouring).
            CB_REQUIRE((size_t)PE * nb2 < 65536, "assembly plan: patch too large for 16-bit entry ids");
            cnt.assign((size_t)h.n_pdof + 1, 0);
            const uint16_t * Lp = &plan.L[(size_t)p * PE * nb2];
            for (int k = 0; k < h.n_elem; ++k)
                for (int a = 0; a < nb2; ++a)
                    if (Lp[entry(k, a)] != 0xFFFF)
                        cnt[Lp[entry(k, a)] + 1]++;
            for (int d = 0; d < h.n_pdof; ++d)
                cnt[d + 1] += cnt[d];
            if (plan.cptr.size() & 1) // keep every patch's offsets 4-byte aligned (copied with 32-bit cp.async)
                plan.cptr.push_back(0);
            h.cptr_begin = (int)plan.cptr.size();
            for (int d = 0; d <= h.n_pdof; ++d)
                plan.cptr.push_back((uint16_t)cnt[d]);
            uint16_t * ce = &plan.cent[(size_t)p * PE * nb2];
            for (int k = 0; k < h.n_elem; ++k)
                for (int a = 0; a < nb2; ++a)
                    if (Lp[entry(k, a)] != 0xFFFF)
                        ce[cnt[Lp[entry(k, a)]]++] = (uint16_t)entry(k, a);
            if (plan.node_major) { // fixed-width records for the helper warps of volume_action_ws
                plan.cent4.resize(plan.gid.size() * 4, 0xFFFF);
                plan.target.resize(plan.gid.size(), 0);
                const uint16_t * cp = &plan.cptr[h.cptr_begin];
                for (int d = 0; d < h.n_pdof; ++d) {
                    uint16_t * r = &plan.cent4[((size_t)h.pdof_begin + d) * 4];
                    const int n = cp[d + 1] - cp[d];
                    for (int k = 0; k < std::min(n, 4); ++k)
                        r[k] = ce[cp[d] + k];
                    if (n > 4)
                        r[3] = 0xFFFE;
                    plan.target[(size_t)h.pdof_begin + d] = d < h.n_int ? plan.gid[(size_t)h.pdof_begin + d] : plan.slot[(size_t)h.slot_begin + d - h.n_int];
                }
            }
        }

    }

    // Host-only self check of an assembly plan (no GPU): plays the kernels' gather / assembly with integer-valued element
    // contributions v(el, a) and compares with the direct sum over I. Exercises both layouts; used by the CPU test-suite.
    // stats: n_patches, PE, listed patch DOFs, shared DOFs, max_pdof, entries beyond four per DOF, mismatches, FNV-1a of all arrays
    void plan_self_check(H1Space & fem, bool tpe, int64_t stats[8])
    {
        Plan plan;
        build_plan(fem, plan, tpe);
        const int nb = fem.nb, nb2 = nb * nb, PE = plan.PE, NG = (nb2 + 3) / 4;
        auto val = [](int64_t el, int a) { return (double)((el * 31 + a * 7) % 1009 + 1); };
        std::vector<double> expect((size_t)fem.ndof, 0.0), y((size_t)fem.ndof, -1.0), partial((size_t)std::max<int64_t>(plan.n_slots_total, 1), -1.0);
        for (int64_t el = 0; el < fem.n_elem; ++el)
            for (int a = 0; a < nb2; ++a)
                expect[fem.I[(size_t)nb2 * el + a]] += val(el, a);
        auto interior = [nb](int a) { const int i = a % nb, j = a / nb; return i > 0 && i < nb - 1 && j > 0 && j < nb - 1; };
        int64_t listed = 0, over4 = 0, bad = 0;
        for (int64_t p = 0; p < plan.n_patches; ++p) {
            const PatchHdr & h = plan.hdr[p];
            auto entry_val = [&](int ent) { // value the compute phase leaves at entry `ent` of the patch buffer
                const int k = plan.node_major ? ent % PE : ent / nb2, a = plan.node_major ? ent / PE : ent % nb2;
                if (k >= h.n_elem)
                    return 1e300; // a padding slot must never be referenced
                return val(plan.slot_elem[(size_t)p * PE + k], a);
            };
            if (plan.node_major) {
                for (int k = 0; k < h.n_elem; ++k) {
                    const int64_t el = plan.slot_elem[(size_t)p * PE + k];
                    for (int a = 0; a < nb2; ++a) {
                        const int gi = plan.Ig[(((size_t)p * NG + a / 4) * PE + k) * 4 + a % 4];
                        if (gi != fem.I[(size_t)nb2 * el + a])
                            ++bad;
                        if (interior(a)) // written by the element thread itself: must be the only contribution
                            y[gi] = (y[gi] == -1.0) ? val(el, a) : 1e300;
                    }
                }
            }
            listed += h.n_pdof;
            const uint16_t * cp = &plan.cptr[h.cptr_begin];
            const uint16_t * ce = &plan.cent[(size_t)p * PE * nb2];
            for (int d = 0; d < h.n_pdof; ++d) {
                double sum = 0.0;
                if (plan.node_major) {
                    const uint16_t * r = &plan.cent4[((size_t)h.pdof_begin + d) * 4];
                    for (int k = 0; k < 3; ++k)
                        if (r[k] != 0xFFFF)
                            sum += entry_val(r[k]);
                    if (r[3] < 0xFFFE)
                        sum += entry_val(r[3]);
                    else if (r[3] == 0xFFFE) {
                        ++over4;
                        for (int k = cp[d] + 3; k < cp[d + 1]; ++k)
                            sum += entry_val(ce[k]);
                    }
                }
                else
                    for (int k = cp[d]; k < cp[d + 1]; ++k)
                        sum += entry_val(ce[k]);
                const bool priv = d < h.n_int;
                const int tgt = plan.node_major ? plan.target[(size_t)h.pdof_begin + d]
                                                : (priv ? plan.gid[(size_t)h.pdof_begin + d] : plan.slot[(size_t)h.slot_begin + d - h.n_int]);
                if (priv)
                    y[tgt] = (y[tgt] == -1.0) ? sum : 1e300;
                else
                    partial[tgt] = (partial[tgt] == -1.0) ? sum : 1e300;
            }
        }
        for (int64_t sidx = 0; sidx < plan.n_shared; ++sidx) {
            double sum = 0.0;
            for (int k = plan.sh_ptr[sidx]; k < plan.sh_ptr[sidx + 1]; ++k)
                sum += partial[k];
            const int gi = plan.sh_gid[sidx];
            y[gi] = (y[gi] == -1.0) ? sum : 1e300;
        }
        for (int64_t g = 0; g < fem.ndof; ++g)
            if (y[g] != expect[g])
                ++bad;
        stats[0] = plan.n_patches;
        stats[1] = PE;
        stats[2] = listed;
        stats[3] = plan.n_shared;
        stats[4] = plan.max_pdof;
        stats[5] = over4;
        stats[6] = bad;
        // FNV-1a over every array of the plan: lets a test pin the plan builder bit for bit
        uint64_t hsh = 0xcbf29ce484222325ull;
        auto mix = [&hsh](const void * ptr, size_t bytes) {
            const unsigned char * b = static_cast<const unsigned char *>(ptr);
            for (size_t i = 0; i < bytes; ++i) {
                hsh ^= b[i];
                hsh *= 0x100000001b3ull;
            }
        };
        auto mixv = [&mix](const auto & v) {
            if (!v.empty())
                mix(v.data(), v.size() * sizeof(v[0]));
        };
        mixv(plan.hdr); mixv(plan.gid); mixv(plan.slot); mixv(plan.L); mixv(plan.cptr); mixv(plan.cent); mixv(plan.slot_elem);
        mixv(plan.Ig); mixv(plan.cent4); mixv(plan.target); mixv(plan.sh_gid); mixv(plan.sh_ptr);
        const int64_t scal[4] = {plan.n_slots_total, plan.n_shared, plan.max_pdof, plan.max_nsh};
        mix(scal, sizeof(scal));
        stats[7] = (int64_t)hsh;
    }

    void Plan::ensure_device()
    {
        if (on_device)
            return;
        d_hdr.upload(hdr);
        d_gid.upload(gid);
        d_slot.upload(slot);
        if (!node_major) // the thread-per-element kernels gather through Ig; L only serves the host-side plan build there
            d_L.upload(L);
        cptr.resize(cptr.size() + 8, 0); // slack: the last patch's offsets are copied in 32-bit words
        d_cptr.upload(cptr);
        d_cent.upload(cent);
        d_slot_elem.upload(slot_elem);
        if (node_major) {
            d_Ig.upload(Ig);
            d_cent4.upload(cent4);
            d_target.upload(target);
        }
        d_sh_gid.upload(sh_gid);
        d_sh_ptr.upload(sh_ptr);
        on_device = true;
    }
} // namespace cb200
