// cuddh_config.hpp — build options of the drop-in header layer (the reference generates this file from config.in).
#ifndef CUDDH_CONFIG_HPP
#define CUDDH_CONFIG_HPP
// #define CUDDH_DEBUG        // bounds / null checks in TensorWrapper
// #define CUDDH_LOG_MEMCPY   // log HostDeviceArray allocations and copies
#define CUDDH_B200 1
#endif
