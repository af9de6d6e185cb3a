// Mesh2D (reference include/Mesh2D.hpp): quadrilateral mesh built by uniform_rect / from_vertices. The topology lives
// in the library (csrc/mesh.cpp, 64-bit edge table, same numbering as the reference); this class is a move-only host
// handle that materialises Edge / Element / Node objects and metric arrays on demand.
#ifndef CUDDH_MESH_2D_HPP
#define CUDDH_MESH_2D_HPP

#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "Edge.hpp"
#include "Element.hpp"
#include "HostDeviceArray.hpp"
#include "Node.hpp"
#include "QuadratureRule.hpp"
#include "Tensor.hpp"

namespace cuddh
{
    class Mesh2D
    {
    public:
        /// metric arrays of all elements on the tensor grid of a 1-D rule (host loop as in source/Mesh2D.cpp:173-227; the
        /// operators of this library do not use these arrays — they evaluate the metrics in their own setup kernels)
        class ElementMetricCollection
        {
        public:
            ElementMetricCollection(const Mesh2D & mesh_, const QuadratureRule & quad_) : mesh(mesh_), quad(quad_) {}
            ElementMetricCollection(ElementMetricCollection && a)
                : mesh(a.mesh), quad(std::move(a.quad)), J(std::move(a.J)), detJ(std::move(a.detJ)), x(std::move(a.x)) {}

            /// (2, 2, n, n, n_elem)
            const double * jacobians(MemorySpace m) const
            {
                if (J.size() == 0)
                    fill(J, 4, [](double * out, const Element * e, const double * xi) { e->jacobian(xi, out); });
                return J.read(m);
            }
            /// (n, n, n_elem)
            const double * measures(MemorySpace m) const
            {
                if (detJ.size() == 0)
                    fill(detJ, 1, [](double * out, const Element * e, const double * xi) { *out = e->measure(xi); });
                return detJ.read(m);
            }
            /// (2, n, n, n_elem)
            const double * physical_coordinates(MemorySpace m) const
            {
                if (x.size() == 0)
                    fill(x, 2, [](double * out, const Element * e, const double * xi) { e->physical_coordinates(xi, out); });
                return x.read(m);
            }

        private:
            template <typename Eval>
            void fill(host_device_dvec & arr, int dim, Eval eval) const
            {
                const int m = quad.size(), nel = mesh.n_elem();
                arr.resize(dim * m * m * nel);
                auto a = reshape(arr.host_write(), dim, m, m, nel);
                double xi[2];
                for (int el = 0; el < nel; ++el) {
                    const Element * elem = mesh.element(el);
                    for (int j = 0; j < m; ++j) {
                        xi[1] = quad.x(j);
                        for (int i = 0; i < m; ++i) {
                            xi[0] = quad.x(i);
                            eval(&a(0, i, j, el), elem, xi);
                        }
                    }
                }
            }

            const Mesh2D & mesh;
            QuadratureRule quad;
            mutable host_device_dvec J, detJ, x;
        };

        /// metric arrays of a set of edges on a 1-D rule: measures (n, ne), coordinates (2, n, ne), normals (2, n, ne)
        class EdgeMetricCollection
        {
        public:
            EdgeMetricCollection(const Mesh2D & mesh_, const FaceType type_, const QuadratureRule & quad_)
                : mesh(mesh_), quad(quad_)
            {
                const int ne = mesh.n_edges(type_);
                ids.resize(ne);
                for (int e = 0; e < ne; ++e)
                    ids[e] = mesh.edge(e, type_)->id;
            }
            EdgeMetricCollection(const Mesh2D & mesh_, int n_faces, const int * faces, const QuadratureRule & quad_)
                : mesh(mesh_), quad(quad_), ids(faces, faces + n_faces) {}
            EdgeMetricCollection(EdgeMetricCollection && a)
                : mesh(a.mesh), quad(std::move(a.quad)), ids(std::move(a.ids)), detJ(std::move(a.detJ)), x(std::move(a.x)), n(std::move(a.n)) {}

            const double * measures(MemorySpace m) const
            {
                if (detJ.size() == 0)
                    fill(detJ, 1, [](double * out, const Edge * E, double xi) { *out = E->measure(xi); });
                return detJ.read(m);
            }
            const double * physical_coordinates(MemorySpace m) const
            {
                if (x.size() == 0)
                    fill(x, 2, [](double * out, const Edge * E, double xi) { E->physical_coordinates(xi, out); });
                return x.read(m);
            }
            const double * normals(MemorySpace m) const
            {
                if (n.size() == 0)
                    fill(n, 2, [](double * out, const Edge * E, double xi) { E->normal(xi, out); });
                return n.read(m);
            }

        private:
            template <typename Eval>
            void fill(host_device_dvec & arr, int dim, Eval eval) const
            {
                const int m = quad.size(), ne = (int)ids.size();
                arr.resize(dim * m * ne);
                auto a = reshape(arr.host_write(), dim, m, ne);
                for (int e = 0; e < ne; ++e) {
                    const Edge * E = mesh.edge(ids[e]);
                    for (int i = 0; i < m; ++i)
                        eval(&a(0, i, e), E, quad.x(i));
                }
            }

            const Mesh2D & mesh;
            QuadratureRule quad;
            std::vector<int> ids;
            mutable host_device_dvec detJ, x, n;
        };

        Mesh2D() {}
        ~Mesh2D() = default;
        Mesh2D(const Mesh2D &) = delete;
        Mesh2D & operator=(const Mesh2D &) = delete;
        Mesh2D(Mesh2D &&) = default;
        Mesh2D & operator=(Mesh2D &&) = default;

        int n_elem() const { return (int)sizes[0]; }
        int n_nodes() const { return (int)sizes[1]; }
        int n_edges() const { return (int)sizes[2]; }
        int n_edges(FaceType type) const { return (int)(type == FaceType::BOUNDARY ? sizes[3] : sizes[4]); }
        int n_nodes(NodeType type) const
        {
            materialise();
            return (int)(type == NodeType::BOUNDARY ? bnodes.size() : inodes.size());
        }
        int max_element_order() const { return 1; } // bilinear elements only
        int min_element_order() const { return 1; }

        double min_h() const { return hmin; }
        double max_h() const { return hmax; }

        const Node & node(int i) const
        {
            materialise();
            return nodes[i];
        }
        const Node & node(int i, NodeType type) const
        {
            materialise();
            return nodes[type == NodeType::BOUNDARY ? bnodes[i] : inodes[i]];
        }
        const Edge * edge(int i) const
        {
            materialise();
            return &edges[i];
        }
        const Edge * edge(int i, FaceType type) const
        {
            materialise();
            return &edges[type == FaceType::BOUNDARY ? bedges[i] : iedges[i]];
        }
        const Element * element(int el) const
        {
            materialise();
            return &elements[el];
        }

        /// indices of the boundary edges, in edge-id order
        ivec boundary_edges() const
        {
            materialise();
            ivec b((int)bedges.size());
            for (size_t i = 0; i < bedges.size(); ++i)
                b((int)i) = bedges[i];
            return b;
        }

        const ElementMetricCollection & element_metrics(const QuadratureRule & quad) const
        {
            const std::string id = quad.name();
            auto it = elem_collections.find(id);
            if (it == elem_collections.end())
                it = elem_collections.emplace(id, ElementMetricCollection(*this, quad)).first;
            return it->second;
        }

        const EdgeMetricCollection & edge_metrics(const QuadratureRule & quad, FaceType type) const
        {
            auto & coll = (type == FaceType::INTERIOR) ? interior_edge_collections : boundary_edge_collections;
            const std::string id = quad.name();
            auto it = coll.find(id);
            if (it == coll.end())
                it = coll.emplace(id, EdgeMetricCollection(*this, type, quad)).first;
            return it->second;
        }

        /// x: (2, nx) vertex coordinates; elems: (4, nel) counter-clockwise corner indices
        static Mesh2D from_vertices(int nx, const double * x, int nel, const int * elems)
        {
            cuddh_mesh_t raw = nullptr;
            cuddh_check(cuddh_b200_mesh_from_vertices(nx, x, nel, elems, &raw));
            return Mesh2D(raw);
        }

        /// nx x ny uniform elements on [ax, bx] x [ay, by]
        static Mesh2D uniform_rect(int nx, double ax, double bx, int ny, double ay, double by)
        {
            cuddh_mesh_t raw = nullptr;
            cuddh_check(cuddh_b200_mesh_uniform_rect(nx, ax, bx, ny, ay, by, &raw));
            return Mesh2D(raw);
        }

        /// library handle (used by H1Space)
        cuddh_mesh_t handle() const { return h.get(); }

    private:
        explicit Mesh2D(cuddh_mesh_t raw) : h(raw, [](cuddh_mesh_t p) { cuddh_b200_mesh_destroy(p); })
        {
            cuddh_check(cuddh_b200_mesh_sizes(raw, sizes));
            cuddh_check(cuddh_b200_mesh_h(raw, &hmin, &hmax));
        }

        // Edge / Element / Node objects are built the first time user code asks for one
        void materialise() const
        {
            if (built)
                return;
            const int ne = n_edges(), nel = n_elem(), nv = n_nodes();
            std::vector<int> rec(8 * (size_t)ne), el(4 * (size_t)nel);
            std::vector<double> xy(2 * (size_t)nv);
            cuddh_check(cuddh_b200_mesh_edges(h.get(), rec.data()));
            cuddh_check(cuddh_b200_mesh_elements(h.get(), el.data()));
            cuddh_check(cuddh_b200_mesh_vertices(h.get(), xy.data()));

            nodes.resize(nv);
            for (int k = 0; k < nv; ++k) {
                nodes[k].id = k;
                nodes[k].type = NodeType::BOUNDARY;
                nodes[k].x[0] = xy[2 * k];
                nodes[k].x[1] = xy[2 * k + 1];
            }
            elements.reserve(nel);
            for (int e = 0; e < nel; ++e) {
                double X[8];
                for (int c = 0; c < 4; ++c) {
                    const int v = el[4 * (size_t)e + c];
                    X[2 * c] = xy[2 * (size_t)v];
                    X[2 * c + 1] = xy[2 * (size_t)v + 1];
                    nodes[v].connected_elements.push_back({c, e});
                }
                elements.emplace_back(X);
                elements.back().id = e;
                for (int c = 0; c < 4; ++c)
                    elements.back().nodes[c] = el[4 * (size_t)e + c];
            }
            edges.reserve(ne);
            for (int e = 0; e < ne; ++e) {
                const int * r = &rec[8 * (size_t)e];
                edges.emplace_back(&xy[2 * (size_t)r[0]], &xy[2 * (size_t)r[1]], r[4]);
                StraightEdge & E = edges.back();
                E.id = e;
                E.nodes[0] = r[0];
                E.nodes[1] = r[1];
                E.elements[0] = r[2];
                E.elements[1] = r[3];
                E.sides[0] = r[4];
                E.sides[1] = r[5];
                E.delta = r[6];
                E.type = r[7] ? FaceType::BOUNDARY : FaceType::INTERIOR;
                if (r[7])
                    bedges.push_back(e);
                else {
                    iedges.push_back(e);
                    nodes[r[0]].type = NodeType::INTERIOR; // as the reference: end points of an interior edge
                    nodes[r[1]].type = NodeType::INTERIOR;
                }
            }
            for (int k = 0; k < nv; ++k)
                (nodes[k].type == NodeType::BOUNDARY ? bnodes : inodes).push_back(k);
            built = true;
        }

        std::shared_ptr<cuddh_mesh_s> h;
        int64_t sizes[5] = {0, 0, 0, 0, 0};
        double hmin = std::numeric_limits<double>::infinity(), hmax = -1;

        mutable bool built = false;
        mutable std::vector<Node> nodes;
        mutable std::vector<StraightEdge> edges;
        mutable std::vector<QuadElement> elements;
        mutable std::vector<int> bedges, iedges, bnodes, inodes;

        mutable std::unordered_map<std::string, ElementMetricCollection> elem_collections;
        mutable std::unordered_map<std::string, EdgeMetricCollection> interior_edge_collections;
        mutable std::unordered_map<std::string, EdgeMetricCollection> boundary_edge_collections;
    };
} // namespace cuddh

#endif
