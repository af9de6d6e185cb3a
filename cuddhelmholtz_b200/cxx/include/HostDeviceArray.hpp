// HostDeviceArray<T>: lazily mirrored host/device buffer with validity flags — the container every user-side vector
// lives in (reference include/HostDeviceArray.hpp). Same methods and copy-on-demand semantics; sizes are 64-bit
// internally (the reference computes byte counts in int and fails at 2 GiB); cudaMalloc failure throws
// std::runtime_error as in the reference, fresh memory is zero on both sides.
#ifndef CUDDH_HOST_DEVICE_ARRAY_HPP
#define CUDDH_HOST_DEVICE_ARRAY_HPP

#include <cstddef>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <utility>

#include <cuda_runtime.h>

#include "cuddh_config.hpp"

namespace cuddh
{
    enum class MemorySpace { HOST, DEVICE };

    template <typename T>
    class HostDeviceArray
    {
    public:
        HostDeviceArray() = default;
        explicit HostDeviceArray(int n_) : n(n_ > 0 ? (size_t)n_ : 0) {}
        HostDeviceArray(const HostDeviceArray &) = delete;
        HostDeviceArray & operator=(const HostDeviceArray &) = delete;
        HostDeviceArray(HostDeviceArray && o) noexcept { steal(o); }
        HostDeviceArray & operator=(HostDeviceArray && o) noexcept
        {
            if (this != &o) {
                drop();
                steal(o);
            }
            return *this;
        }
        ~HostDeviceArray() { drop(); }

        int size() const { return (int)n; }

        /// new size; previous contents (both sides) are discarded
        void resize(int new_size)
        {
            drop();
            n = new_size > 0 ? (size_t)new_size : 0;
        }

        const T * read(MemorySpace m) const { return m == MemorySpace::HOST ? host_read() : device_read(); }
        T * write(MemorySpace m) { return m == MemorySpace::HOST ? host_write() : device_write(); }
        T * read_write(MemorySpace m) { return m == MemorySpace::HOST ? host_read_write() : device_read_write(); }

        const T * host_read(bool force_copy = false) const
        {
            if (n == 0) return nullptr;
            if (!host_ok || force_copy) {
                ensure_host();
                if (dev_ok)
                    cudaMemcpy(h, d, n * sizeof(T), cudaMemcpyDeviceToHost);
            }
            host_ok = true;
            return h;
        }
        T * host_write()
        {
            if (n == 0) return nullptr;
            ensure_host();
            host_ok = true;
            dev_ok = false;
            return h;
        }
        T * host_read_write(bool force_copy = false)
        {
            host_read(force_copy);
            return host_write();
        }
        T * host_release()
        {
            T * p = h;
            h = nullptr;
            return p;
        }

        const T * device_read(bool force_copy = false) const
        {
            if (n == 0) return nullptr;
            if (!dev_ok || force_copy) {
                ensure_device();
                if (host_ok)
                    cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice);
            }
            dev_ok = true;
            return d;
        }
        T * device_write()
        {
            if (n == 0) return nullptr;
            ensure_device();
            dev_ok = true;
            host_ok = false;
            return d;
        }
        T * device_read_write(bool force_copy = false)
        {
            device_read(force_copy);
            return device_write();
        }
        T * device_release()
        {
            T * p = d;
            d = nullptr;
            return p;
        }

    private:
        size_t n = 0;
        mutable bool dev_ok = false, host_ok = false;
        mutable T * d = nullptr;
        mutable T * h = nullptr;

        void ensure_host() const
        {
            if (!h) {
#ifdef CUDDH_LOG_MEMCPY
                std::cout << "HostDeviceArray: new host array (" << n * sizeof(T) << " bytes)" << std::endl;
#endif
                h = new T[n]();
            }
        }
        void ensure_device() const
        {
            if (!d) {
#ifdef CUDDH_LOG_MEMCPY
                std::cout << "HostDeviceArray: new device array (" << n * sizeof(T) << " bytes)" << std::endl;
#endif
                const cudaError_t err = cudaMalloc((void **)&d, n * sizeof(T));
                if (err != cudaSuccess)
                    throw std::runtime_error(cudaGetErrorString(err));
                cudaMemset(d, 0, n * sizeof(T));
            }
        }
        void drop()
        {
            delete[] h;
            if (d) cudaFree(d);
            h = nullptr;
            d = nullptr;
            host_ok = dev_ok = false;
        }
        void steal(HostDeviceArray & o)
        {
            n = o.n;
            dev_ok = std::exchange(o.dev_ok, false);
            host_ok = std::exchange(o.host_ok, false);
            d = std::exchange(o.d, nullptr);
            h = std::exchange(o.h, nullptr);
        }
    };

    typedef HostDeviceArray<double> host_device_dvec;
    typedef HostDeviceArray<int> host_device_ivec;
} // namespace cuddh

#endif
