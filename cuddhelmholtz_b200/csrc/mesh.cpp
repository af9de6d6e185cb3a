// Quad-mesh topology with the reference's numbering semantics (source/Mesh2D.cpp:11-171):
//   * edges are discovered in element order, local sides s = 0..3 joining corners (0,1),(1,2),(3,2),(0,3);
//     the first element to touch an edge is elements[0] and fixes the edge id; `delta` is the relative
//     orientation seen by the second element;
//   * boundary / interior lists are in edge-id order.
// Differences by design: flat SoA storage instead of one heap object per edge/element, and a 64-bit
// open-addressing edge table (the reference's 32-bit key `min + nv*max` overflows and corrupts
// uniform_rect(2048)).
#include "common.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>

namespace cb200
{
    namespace
    {
        struct EdgeTable
        {
            std::vector<uint64_t> keys;
            std::vector<int> vals;
            uint64_t mask;
            explicit EdgeTable(size_t expected)
            {
                size_t cap = 16;
                while (cap < 2 * expected + 16)
                    cap <<= 1;
                keys.assign(cap, ~0ull);
                vals.assign(cap, -1);
                mask = cap - 1;
            }
            static uint64_t mix(uint64_t k)
            {
                k ^= k >> 33;
                k *= 0xff51afd7ed558ccdull;
                k ^= k >> 33;
                k *= 0xc4ceb9fe1a85ec53ull;
                k ^= k >> 33;
                return k;
            }
            // returns reference to the value slot for key (inserting with -1 if absent)
            int & slot(uint64_t key)
            {
                uint64_t h = mix(key) & mask;
                while (keys[h] != ~0ull && keys[h] != key)
                    h = (h + 1) & mask;
                keys[h] = key;
                return vals[h];
            }
        };
    } // namespace

    std::unique_ptr<Mesh> Mesh::from_vertices(int64_t nv, const double * xy_, int64_t nel, const int * elems_)
    {
        CB_REQUIRE(nv > 0 && nel > 0, "Mesh2D::from_vertices: empty mesh");
        CB_REQUIRE(nv < (int64_t)std::numeric_limits<int>::max() && 4 * nel < (int64_t)std::numeric_limits<int>::max(),
                   "Mesh2D::from_vertices: mesh too large for 32-bit vertex / edge ids");
        std::unique_ptr<Mesh> m(new Mesh);
        m->n_nodes = nv;
        m->n_elem = nel;
        m->xy.assign(xy_, xy_ + 2 * nv);
        m->elems.assign(elems_, elems_ + 4 * nel);
        for (int64_t i = 0; i < 4 * nel; ++i)
            CB_REQUIRE(elems_[i] >= 0 && elems_[i] < nv, "Mesh2D::from_vertices: element corner index out of range");
        m->elem_edges.assign(4 * (size_t)nel, -1);

        static const int side_a[4] = {0, 1, 3, 0};
        static const int side_b[4] = {1, 2, 2, 3};

        EdgeTable table(2 * (size_t)nel + (size_t)nv);
        std::vector<int> & E = m->edges;
        E.reserve(8 * (2 * (size_t)nel + 1024));
        int n_edges = 0;
        for (int64_t el = 0; el < nel; ++el) {
            for (int s = 0; s < 4; ++s) {
                const int c0 = elems_[4 * el + side_a[s]];
                const int c1 = elems_[4 * el + side_b[s]];
                const uint64_t lo = (uint64_t)std::min(c0, c1), hi = (uint64_t)std::max(c0, c1);
                int & e = table.slot((hi << 32) | lo);
                if (e < 0) { // first touch: this element owns the edge
                    e = n_edges++;
                    const int rec[8] = {c0, c1, (int)el, -1, s, -1, 1, 1};
                    E.insert(E.end(), rec, rec + 8);
                }
                else { // second touch: interior edge
                    int * rec = &E[8 * (size_t)e];
                    CB_REQUIRE(rec[3] < 0, "Mesh2D::from_vertices: an edge is shared by more than two elements");
                    const int first_start = elems_[4 * (int64_t)rec[2] + side_a[rec[4]]];
                    rec[3] = (int)el;
                    rec[5] = s;
                    rec[6] = (c0 == first_start) ? 1 : -1;
                    rec[7] = 0;
                }
                m->elem_edges[4 * el + s] = e;
            }
        }
        m->n_edges = n_edges;
        m->edge_meas.resize(n_edges);
        double hmin = std::numeric_limits<double>::infinity(), hmax = -1;
        for (int e = 0; e < n_edges; ++e) {
            const int * rec = &E[8 * (size_t)e];
            (rec[7] ? m->boundary_edges : m->interior_edges).push_back(e);
            const double dx = xy_[2 * (size_t)rec[1]] - xy_[2 * (size_t)rec[0]];
            const double dy = xy_[2 * (size_t)rec[1] + 1] - xy_[2 * (size_t)rec[0] + 1];
            const double meas = std::hypot(dx, dy) / 2; // StraightEdge: include/Edge.hpp:95-113
            m->edge_meas[e] = meas;
            hmin = std::min(hmin, 2.0 * meas);
            hmax = std::max(hmax, 2.0 * meas);
        }
        m->min_h = hmin;
        m->max_h = hmax;
        return m;
    }

    std::unique_ptr<Mesh> Mesh::uniform_rect(int nx, double ax, double bx, int ny, double ay, double by)
    {
        CB_REQUIRE(nx > 0 && ny > 0, "Mesh2D::uniform_rect: nx, ny must be positive");
        const int64_t np = (int64_t)(nx + 1) * (ny + 1), nel = (int64_t)nx * ny;
        std::vector<double> coo(2 * (size_t)np);
        std::vector<int> el(4 * (size_t)nel);
        const double dx = (bx - ax) / nx, dy = (by - ay) / ny; // vertex formula of source/Mesh2D.cpp:147-157
        for (int j = 0; j <= ny; ++j) {
            const double y = ay + dy * j;
            for (int i = 0; i <= nx; ++i) {
                const size_t v = (size_t)i + (size_t)(nx + 1) * j;
                coo[2 * v] = ax + dx * i;
                coo[2 * v + 1] = y;
            }
        }
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const size_t e = (size_t)i + (size_t)nx * j;
                const int v00 = i + (nx + 1) * j;
                el[4 * e + 0] = v00;
                el[4 * e + 1] = v00 + 1;
                el[4 * e + 2] = v00 + 1 + (nx + 1);
                el[4 * e + 3] = v00 + (nx + 1);
            }
        if (getenv("CUDDH_B200_CLOSED_FORM") && atoi(getenv("CUDDH_B200_CLOSED_FORM")) == 0) { // the generic edge-table path
            auto m = from_vertices(np, coo.data(), nel, el.data());
            m->nx = nx;
            m->ny = ny;
            return m;
        }
        // Closed form of from_vertices() on this vertex / element layout (bit-identical arrays, checked by tests/test_host_setup.py):
        // elements are visited in order el = i + nx j, sides 0..3 = bottom, right, top, left; an element FIRST touches its right and
        // top edge always, its bottom edge only in the first row and its left edge only in the first column, so the edge ids are a
        // prefix sum, every interior edge has delta = +1 (both elements traverse it in the same direction), and no edge table is
        // needed. Runs on all host threads.
        CB_REQUIRE(np < (int64_t)std::numeric_limits<int>::max() && 4 * nel < (int64_t)std::numeric_limits<int>::max(),
                   "Mesh2D::from_vertices: mesh too large for 32-bit vertex / edge ids");
        std::unique_ptr<Mesh> m(new Mesh);
        m->n_nodes = np;
        m->n_elem = nel;
        m->nx = nx;
        m->ny = ny;
        m->xy = std::move(coo);
        m->elems = std::move(el);
        const int64_t n_edges = (int64_t)nx * (ny + 1) + (int64_t)ny * (nx + 1);
        m->n_edges = n_edges;
        // first edge id of every element: 2 + (first row) + (first column) new edges per element
        auto first_id = [nx](int64_t i, int64_t j) -> int64_t {
            // elements before (i, j): j full rows + i elements of row j
            const int64_t rows = j, before = rows * nx + i;
            int64_t id = 2 * before;
            id += (j == 0) ? i : nx;            // bottom edges: only elements of the first row own one
            id += rows + ((i > 0) ? 1 : 0);     // left edges: the first element of every row owns one
            return id;
        };
        m->edges.assign(8 * (size_t)n_edges, 0);
        m->elem_edges.assign(4 * (size_t)nel, -1);
        m->edge_meas.resize((size_t)n_edges);
        int * E = m->edges.data();
        const int * EL = m->elems.data();
        const double * XY = m->xy.data();
        static const int side_a[4] = {0, 1, 3, 0};
        static const int side_b[4] = {1, 2, 2, 3};
        parallel_for(ny, [&](int64_t jb, int64_t je, int) {
            for (int64_t j = jb; j < je; ++j)
                for (int64_t i = 0; i < nx; ++i) {
                    const int64_t e_id = i + (int64_t)nx * j;
                    const int64_t base = first_id(i, j);
                    const int own0 = (j == 0) ? 1 : 0, own3 = (i == 0) ? 1 : 0;
                    const int64_t id_of[4] = {own0 ? base : first_id(i, j - 1) + ((j - 1 == 0) ? 1 : 0) + 1, // top edge of the element below
                                              base + own0, base + own0 + 1,
                                              own3 ? base + own0 + 2 : first_id(i - 1, j) + ((j == 0) ? 1 : 0)}; // right edge of the left neighbour
                    for (int s = 0; s < 4; ++s) {
                        m->elem_edges[4 * e_id + s] = (int)id_of[s];
                        const bool own = (s == 1 || s == 2) || (s == 0 && own0) || (s == 3 && own3);
                        if (!own)
                            continue;
                        int * rec = E + 8 * id_of[s];
                        const int c0 = EL[4 * e_id + side_a[s]], c1 = EL[4 * e_id + side_b[s]];
                        rec[0] = c0;
                        rec[1] = c1;
                        rec[2] = (int)e_id;
                        rec[4] = s;
                        rec[6] = 1;
                        // second element: the right edge is the left side of (i+1, j), the top edge the bottom side of (i, j+1)
                        const bool interior = (s == 1 && i + 1 < nx) || (s == 2 && j + 1 < ny);
                        rec[3] = interior ? (int)(s == 1 ? e_id + 1 : e_id + nx) : -1;
                        rec[5] = interior ? (s == 1 ? 3 : 0) : -1;
                        rec[7] = interior ? 0 : 1;
                        const double ddx = XY[2 * (size_t)c1] - XY[2 * (size_t)c0], ddy = XY[2 * (size_t)c1 + 1] - XY[2 * (size_t)c0 + 1];
                        m->edge_meas[(size_t)id_of[s]] = std::hypot(ddx, ddy) / 2; // StraightEdge: include/Edge.hpp:95-113
                    }
                }
        });
        double hmin = std::numeric_limits<double>::infinity(), hmax = -1;
        m->boundary_edges.reserve(2 * ((size_t)nx + ny));
        m->interior_edges.reserve((size_t)n_edges);
        for (int64_t e = 0; e < n_edges; ++e) {
            (E[8 * e + 7] ? m->boundary_edges : m->interior_edges).push_back((int)e);
            hmin = std::min(hmin, 2.0 * m->edge_meas[(size_t)e]);
            hmax = std::max(hmax, 2.0 * m->edge_meas[(size_t)e]);
        }
        m->min_h = hmin;
        m->max_h = hmax;
        return m;
    }
} // namespace cb200
