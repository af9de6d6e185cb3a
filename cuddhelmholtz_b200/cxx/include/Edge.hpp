// Edge / StraightEdge (reference include/Edge.hpp:9-157): topology record of a mesh edge plus the straight-segment map.
#ifndef CUDDH_EDGE_HPP
#define CUDDH_EDGE_HPP

#include <cmath>

#include "cuddh_config.hpp"

namespace cuddh
{
    enum class FaceType
    {
        INTERIOR, ///< two elements: elements[1] / sides[1] are defined
        BOUNDARY  ///< one element only
    };

    struct Edge
    {
        FaceType type;
        int id;           ///< global edge index
        int nodes[2];     ///< end points
        int elements[2];  ///< elements[0] = first element that touched the edge
        int sides[2];     ///< local side (0..3) in each element
        int delta;        ///< +1 / -1: orientation seen from the second element

        virtual void normal(const double xi, double * n) const = 0;
        virtual double measure(const double xi) const = 0;
        virtual void physical_coordinates(const double xi, double * x) const = 0;
        virtual double length() const = 0;

        Edge() : id{-1}, elements{-1, -1}, sides{-1, -1} {}
        virtual ~Edge() = default;
    };

    // A straight segment between two mesh vertices. Only the segment itself is stored (start point, end - start, its length and
    // whether the outward normal of the first element is the left or the right normal of the direction of travel); normal,
    // measure and coordinates are evaluated from that on request - the values the reference's StraightEdge returns
    // (include/Edge.hpp:95-157), which is also what the C ABI's edge records carry (cuddh_b200_mesh_edges: end points, first
    // element, side).
    struct StraightEdge : public Edge
    {
    private:
        double a_[2];  // start point
        double d_[2];  // end point - start point
        double len_;   // |d_|
        bool flip_;    // sides 2 and 3 of the first element are traversed against the element's orientation

    public:
        StraightEdge(const double * x0, const double * x1, int side)
            : a_{x0[0], x0[1]}, d_{x1[0] - x0[0], x1[1] - x0[1]}, len_(std::hypot(x1[0] - x0[0], x1[1] - x0[1])), flip_(side == 2 || side == 3)
        {
        }

        void normal(const double, double * n_) const override // d_ rotated by -90 degrees, unit length, flipped for sides 2 / 3
        {
            const double o = flip_ ? -1.0 : 1.0;
            n_[0] = o * d_[1] / len_;
            n_[1] = -o * d_[0] / len_;
        }
        double measure(const double) const override { return len_ / 2; } // d(x)/d(xi) on xi in [-1, 1]
        void physical_coordinates(const double xi, double * x_) const override
        {
            const double t = 0.5 * (xi + 1.0);
            x_[0] = a_[0] + d_[0] * t;
            x_[1] = a_[1] + d_[1] * t;
        }
        double length() const override { return len_; }
    };
} // namespace cuddh

#endif
