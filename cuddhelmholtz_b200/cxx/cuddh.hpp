// cuddh.hpp — umbrella header of the B200 drop-in for CuDDHelmholtz (same name and contents as the reference's
// cuddh.hpp:4-25). User code keeps `#include "cuddh.hpp"`, namespace cuddh, and links libcuddh_b200.so instead of
// libcuddh.a. Every class below is a thin C++ host object over the C ABI in include/cuddh_b200.h.
#ifndef CUDDH_HPP
#define CUDDH_HPP

#include "cuddh_config.hpp"

#include "include/Basis.hpp"
#include "include/cuddh_error.hpp"
#include "include/DDH.hpp"
#include "include/Edge.hpp"
#include "include/Element.hpp"
#include "include/EnsembleSpace.hpp"
#include "include/FaceLinearFunctional.hpp"
#include "include/FaceMassMatrix.hpp"
#include "include/forall.hpp"
#include "include/gmres.hpp"
#include "include/H1Space.hpp"
#include "include/HostDeviceArray.hpp"
#include "include/linalg.hpp"
#include "include/LinearFunctional.hpp"
#include "include/MassMatrix.hpp"
#include "include/Mesh2D.hpp"
#include "include/Node.hpp"
#include "include/Operator.hpp"
#include "include/QuadratureRule.hpp"
#include "include/StiffnessMatrix.hpp"
#include "include/Tensor.hpp"

#endif
