// Multi-GPU plumbing below the C ABI (new: the reference is single-GPU). One process per GPU; NCCL is
// reached through dlopen (no link-time dependency: the library loads and every single-GPU path works on a box without
// NCCL, and inside a torch process it binds to the NCCL copy torch already loaded).
#pragma once
#include "common.hpp"

namespace cb200
{
    struct Comm
    {
        void * nccl = nullptr; // ncclComm_t
        int rank = 0, world = 1;
        bool owned = false;    // created by comm_create (destroyed with the handle) or wrapped (caller-owned)
        ~Comm();
    };

    // 128-byte ncclUniqueId, made on one rank and distributed by the caller's bootstrap (MPI, torch.distributed, a file)
    void comm_unique_id(unsigned char id[128]);
    std::unique_ptr<Comm> comm_create(const unsigned char id[128], int rank, int world);
    std::unique_ptr<Comm> comm_wrap(void * nccl_comm, int rank, int world);

    // in-place sum over ranks, stream-ordered (comm == null or world == 1: no-op)
    void comm_allreduce_sum(const Comm * c, double * d_buf, int64_t count, cudaStream_t s);
    void comm_allreduce_sum(const Comm * c, float * d_buf, int64_t count, cudaStream_t s);

    // one grouped exchange: for every peer p, send send_count[p] floats from d_send + send_off[p] and receive
    // recv_count[p] floats into d_recv + recv_off[p] (zero counts are skipped)
    struct PeerSeg
    {
        int peer;
        int64_t send_off, send_count, recv_off, recv_count;
    };
    void comm_exchange(const Comm * c, const std::vector<PeerSeg> & segs, const void * d_send, void * d_recv, size_t elem_size, cudaStream_t s);
    bool nccl_available();
} // namespace cb200
