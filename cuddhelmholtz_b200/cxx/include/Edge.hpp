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

    struct StraightEdge : public Edge
    {
    private:
        double n[2];
        double meas;
        double x[2];
        double dx[2];

    public:
        /// end points x0 -> x1; `side` of the first element fixes the sign of the outward normal
        StraightEdge(const double * x0, const double * x1, int side)
        {
            x[0] = x0[0];
            x[1] = x0[1];
            dx[0] = x1[0] - x0[0];
            dx[1] = x1[1] - x0[1];
            const double s = std::hypot(dx[0], dx[1]);
            const double sgn = (side == 2 || side == 3) ? -1 : 1;
            n[0] = sgn * dx[1] / s;
            n[1] = -sgn * dx[0] / s;
            meas = s / 2;
        }

        void normal(const double, double * n_) const override
        {
            n_[0] = n[0];
            n_[1] = n[1];
        }
        double measure(const double) const override { return meas; }
        void physical_coordinates(const double xi, double * x_) const override
        {
            const double t = 0.5 * (xi + 1.0);
            x_[0] = x[0] + dx[0] * t;
            x_[1] = x[1] + dx[1] * t;
        }
        double length() const override { return 2.0 * meas; }
    };
} // namespace cuddh

#endif
