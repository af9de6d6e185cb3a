// NCCL through dlopen: see dist.hpp. Only the handful of entry points the solve path needs (neighbour
// send/recv of interface traces, allreduce of the Krylov inner products).
#include "dist.hpp"
#include <dlfcn.h>
#include <mutex>

namespace cb200
{
    namespace
    {
        struct UniqueId
        {
            char internal[128];
        };
        // nccl.h enums (stable across NCCL 2.x)
        constexpr int kNcclUint8 = 1, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

        struct Api
        {
            void * lib = nullptr;
            int (*GetUniqueId)(UniqueId *) = nullptr;
            int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
            int (*CommDestroy)(void *) = nullptr;
            int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
            int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
            int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
            int (*GroupStart)() = nullptr;
            int (*GroupEnd)() = nullptr;
            const char * (*GetErrorString)(int) = nullptr;
            bool ok = false;
        };

        Api & api()
        {
            static Api a;
            static std::once_flag once;
            std::call_once(once, [] {
                // the copy already in the process (torch's bundled NCCL) first, then the system one
                const char * names[] = {"libnccl.so.2", "libnccl.so"};
                for (const char * n : names) {
                    a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
                    if (a.lib)
                        break;
                }
                for (const char * n : names) {
                    if (a.lib)
                        break;
                    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                }
                if (!a.lib)
                    return;
                auto sym = [&](const char * s) { return dlsym(a.lib, s); };
                a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
                a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
                a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
                a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
                a.Send = (decltype(a.Send))sym("ncclSend");
                a.Recv = (decltype(a.Recv))sym("ncclRecv");
                a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
                a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
                a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
                a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
            });
            return a;
        }

        Api & need_api()
        {
            Api & a = api();
            if (!a.ok)
                throw Error(-3, "NCCL is not available in this process (libnccl.so.2 could not be loaded)");
            return a;
        }

        void nccl_check(int rc, const char * what)
        {
            if (rc != 0) {
                Api & a = api();
                throw Error(1000 + rc, std::string(what) + " failed: " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
            }
        }
    } // namespace

    bool nccl_available() { return api().ok; }

    Comm::~Comm()
    {
        if (owned && nccl && api().ok)
            api().CommDestroy(nccl);
    }

    void comm_unique_id(unsigned char id[128])
    {
        UniqueId u;
        nccl_check(need_api().GetUniqueId(&u), "ncclGetUniqueId");
        std::memcpy(id, u.internal, 128);
    }

    std::unique_ptr<Comm> comm_create(const unsigned char id[128], int rank, int world)
    {
        CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_create: rank out of range");
        std::unique_ptr<Comm> c(new Comm);
        c->rank = rank;
        c->world = world;
        UniqueId u;
        std::memcpy(u.internal, id, 128);
        nccl_check(need_api().CommInitRank(&c->nccl, world, u, rank), "ncclCommInitRank");
        c->owned = true;
        return c;
    }

    std::unique_ptr<Comm> comm_wrap(void * nccl_comm, int rank, int world)
    {
        CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_wrap: rank out of range");
        CB_REQUIRE(nccl_comm != nullptr || world == 1, "comm_wrap: null ncclComm_t");
        if (world > 1)
            need_api();
        std::unique_ptr<Comm> c(new Comm);
        c->nccl = nccl_comm;
        c->rank = rank;
        c->world = world;
        c->owned = false;
        return c;
    }

    void comm_allreduce_sum(const Comm * c, double * d_buf, int64_t count, cudaStream_t s)
    {
        if (!c || c->world == 1 || count <= 0)
            return;
        nccl_check(need_api().AllReduce(d_buf, d_buf, (size_t)count, kNcclFloat64, kNcclSum, c->nccl, s), "ncclAllReduce");
    }
    void comm_allreduce_sum(const Comm * c, float * d_buf, int64_t count, cudaStream_t s)
    {
        if (!c || c->world == 1 || count <= 0)
            return;
        nccl_check(need_api().AllReduce(d_buf, d_buf, (size_t)count, kNcclFloat32, kNcclSum, c->nccl, s), "ncclAllReduce");
    }

    void comm_exchange(const Comm * c, const std::vector<PeerSeg> & segs, const void * d_send, void * d_recv, size_t elem_size, cudaStream_t s)
    {
        if (!c || c->world == 1 || segs.empty())
            return;
        Api & a = need_api();
        nccl_check(a.GroupStart(), "ncclGroupStart");
        for (const PeerSeg & g : segs) {
            // bytes on the wire: the element type does not matter to a copy
            if (g.send_count > 0)
                nccl_check(a.Send((const char *)d_send + g.send_off * elem_size, (size_t)g.send_count * elem_size, kNcclUint8, g.peer, c->nccl, s),
                           "ncclSend");
            if (g.recv_count > 0)
                nccl_check(a.Recv((char *)d_recv + g.recv_off * elem_size, (size_t)g.recv_count * elem_size, kNcclUint8, g.peer, c->nccl, s), "ncclRecv");
        }
        nccl_check(a.GroupEnd(), "ncclGroupEnd");
    }
} // namespace cb200
