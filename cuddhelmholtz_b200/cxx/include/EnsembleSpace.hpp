// EnsembleSpace: index maps of a decomposition of the H1 space into labelled subspaces
// (reference include/EnsembleSpace.hpp:13-140). Built by the library (csrc/ddh_setup.cpp) with the reference's
// numbering; arrays are exposed through the same accessors, with DEVICE copies made on first request.
#ifndef CUDDH_ENSEMBLE_SPACE_HPP
#define CUDDH_ENSEMBLE_SPACE_HPP

#include <algorithm>
#include <array>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>

#include "H1Space.hpp"

namespace cuddh
{
    class EnsembleSpace
    {
    public:
        /// element_labels[el] in [0, n_spaces): the subspace element el belongs to
        EnsembleSpace(const H1Space & fem, int n_spaces_, const int * element_labels) : n_spaces(n_spaces_), n_basis(fem.basis().size())
        {
            cuddh_ensemble_t raw = nullptr;
            cuddh_check(cuddh_b200_ensemble_create(fem.handle(), n_spaces, element_labels, &raw));
            h.reset(raw, [](cuddh_ensemble_t p) { cuddh_b200_ensemble_destroy(p); });
            int64_t info[6];
            cuddh_check(cuddh_b200_ensemble_info(raw, info));
            mx_elems = (int)info[1];
            mx_faces = (int)info[2];
            mx_ndof = (int)info[3];
            mx_fdof = (int)info[4];
            n_shared_dofs = (int)info[5];
        }

        int size() const { return n_spaces; }

        const_imat_wrapper global_indices(MemorySpace m) const { return reshape(arr("gI", m), mx_ndof, n_spaces); }
        const_ivec_wrapper sizes(MemorySpace m) const { return reshape(arr("sizes", m), n_spaces); }
        const_imat_wrapper elements(MemorySpace m) const { return reshape(arr("elements", m), mx_elems, n_spaces); }
        const_ivec_wrapper n_elems(MemorySpace m) const { return reshape(arr("n_elems", m), n_spaces); }
        const_imat_wrapper faces(MemorySpace m) const { return reshape(arr("faces", m), mx_faces, n_spaces); }
        const_ivec_wrapper n_faces(MemorySpace m) const { return reshape(arr("n_faces", m), n_spaces); }
        TensorWrapper<4, const int> subspace_indices(MemorySpace m) const { return reshape(arr("sI", m), n_basis, n_basis, mx_elems, n_spaces); }
        const_icube_wrapper face_indices(MemorySpace m) const { return reshape(arr("fI", m), n_basis, mx_faces, n_spaces); }
        const_imat_wrapper face_proj(MemorySpace m) const { return reshape(arr("pI", m), mx_fdof, n_spaces); }
        const_ivec_wrapper fsizes(MemorySpace m) const { return reshape(arr("fsizes", m), n_spaces); }
        /// (4, n_shared): [p, q, i, j] — face DOF i of subspace p is face DOF j of subspace q
        const_imat_wrapper connectivity_map(MemorySpace m) const { return reshape(arr("cmap", m), 4, n_shared_dofs); }

    private:
        const int * arr(const char * name, MemorySpace m) const
        {
            int64_t count = 0;
            const int * host = cuddh_b200_ensemble_array(h.get(), name, &count);
            if (!host)
                cuddh_error(cuddh_b200_last_error());
            if (m == MemorySpace::HOST)
                return host;
            auto it = dev.find(name);
            if (it == dev.end()) {
                it = dev.emplace(name, host_device_ivec((int)count)).first;
                std::copy(host, host + count, it->second.host_write());
            }
            return it->second.device_read();
        }

        const int n_spaces;
        const int n_basis;
        int mx_elems, mx_faces, mx_ndof, mx_fdof, n_shared_dofs;
        std::shared_ptr<cuddh_ensemble_s> h;
        mutable std::unordered_map<std::string, host_device_ivec> dev;
    };
} // namespace cuddh

#endif
