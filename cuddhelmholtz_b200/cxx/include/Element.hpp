// Element / QuadElement (reference include/Element.hpp:9-120, source/Element.cpp): bilinear quadrilateral map.
#ifndef CUDDH_ELEMENT_HPP
#define CUDDH_ELEMENT_HPP

#include "Tensor.hpp"
#include "cuddh_config.hpp"

namespace cuddh
{
    class Element
    {
    public:
        int id;
        int nodes[4];

        virtual void physical_coordinates(const double * xi, double * x) const = 0;
        virtual void jacobian(const double * xi, double * J) const = 0; ///< J = {x_xi, y_xi, x_eta, y_eta}
        virtual double measure(const double * xi) const
        {
            double J[4];
            jacobian(xi, J);
            return J[0] * J[3] - J[1] * J[2];
        }
        virtual double area() const = 0;
        virtual ~Element() = default;
    };

    class QuadElement : public Element
    {
        double x[4][2];

    public:
        /// X: (2, 4) column-major corner coordinates, counter-clockwise
        explicit QuadElement(const double * X)
        {
            for (int i = 0; i < 4; ++i) {
                x[i][0] = X[2 * i];
                x[i][1] = X[2 * i + 1];
            }
        }

        void physical_coordinates(const double * xi, double * x_) const override
        {
            const double b[] = {0.25 * (1.0 - xi[0]) * (1.0 - xi[1]), 0.25 * (1.0 + xi[0]) * (1.0 - xi[1]),
                                0.25 * (1.0 + xi[0]) * (1.0 + xi[1]), 0.25 * (1.0 - xi[0]) * (1.0 + xi[1])};
            x_[0] = 0.0;
            x_[1] = 0.0;
            for (int i = 0; i < 4; ++i) {
                x_[0] += x[i][0] * b[i];
                x_[1] += x[i][1] * b[i];
            }
        }

        void jacobian(const double * xi, double * J) const override
        {
            J[0] = 0.25 * ((1.0 - xi[1]) * (x[1][0] - x[0][0]) + (1.0 + xi[1]) * (x[2][0] - x[3][0]));
            J[1] = 0.25 * ((1.0 - xi[1]) * (x[1][1] - x[0][1]) + (1.0 + xi[1]) * (x[2][1] - x[3][1]));
            J[2] = 0.25 * ((1.0 - xi[0]) * (x[3][0] - x[0][0]) + (1.0 + xi[0]) * (x[2][0] - x[1][0]));
            J[3] = 0.25 * ((1.0 - xi[0]) * (x[3][1] - x[0][1]) + (1.0 + xi[0]) * (x[2][1] - x[1][1]));
        }

        double area() const override
        {
            const double zero[] = {0.0, 0.0};
            return 4.0 * measure(zero); // det J is bilinear: one-point Gauss rule is exact
        }

        const double * corner(int i) const { return x[i]; }
    };
} // namespace cuddh

#endif
