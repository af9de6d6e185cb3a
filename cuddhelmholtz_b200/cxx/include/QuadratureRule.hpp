// QuadratureRule(n, type): Gauss-Legendre / Gauss-Lobatto rule on [-1,1] (reference include/QuadratureRule.hpp:13-77).
// Nodes and weights come from the library (csrc/tables.cpp), which follows the reference's formulas.
#ifndef CUDDH_QUADRATURE_HPP
#define CUDDH_QUADRATURE_HPP

#include <cmath>
#include <iomanip>
#include <sstream>
#include <string>

#include "Tensor.hpp"
#include "cuddh_error.hpp"

namespace cuddh
{
    class QuadratureRule
    {
    public:
        enum QuadratureType { GaussLegendre, GaussLobatto };

        QuadratureRule() : _n(0), _type(GaussLobatto) {}
        QuadratureRule(int n, QuadratureType type = GaussLobatto) : _n(n), _type(type), _x(n), _w(n)
        {
            cuddh_check(cuddh_b200_quadrature(n, type == GaussLegendre ? CUDDH_GAUSS_LEGENDRE : CUDDH_GAUSS_LOBATTO, _x.data(), _w.data()));
        }

        int size() const { return _n; }
        QuadratureType type() const { return _type; }

        /// "%s%05d": "legendre" / "lobatto" followed by n
        std::string name() const
        {
            std::stringstream s;
            s << (_type == GaussLegendre ? "legendre" : "lobatto") << std::setw(5) << std::setfill('0') << _n;
            return s.str();
        }

        const_dvec_wrapper x() const { return const_dvec_wrapper(_x.data(), _n); }
        double x(int i) const { return _x(i); }
        const_dvec_wrapper w() const { return const_dvec_wrapper(_w.data(), _n); }
        double w(int i) const { return _w(i); }

    private:
        int _n;
        QuadratureType _type;
        dvec _x, _w;
    };
} // namespace cuddh

#endif
