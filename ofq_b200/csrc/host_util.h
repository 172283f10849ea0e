// Host-side helpers shared by the C-ABI translation units: error reporting, arch gate, TMA map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "ofq_b200.h"

void ofq_set_error(const char* fmt, ...);
int ofq_check_arch();  // 0 if current device is sm_100, else OFQ_ERR_ARCH (message set)
int ofq_encode_tensor_map(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, void* addr,
                          const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                          const cuuint32_t* estr);
int ofq_encode_tensor_map_sw(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, void* addr,
                             const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                             const cuuint32_t* estr, CUtensorMapSwizzle swizzle);
int ofq_num_sms();

#define OFQ_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ofq_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                          __LINE__);                                                         \
            return OFQ_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define OFQ_CHECK_ARCH()                 \
    do {                                 \
        int _a = ofq_check_arch();       \
        if (_a) return _a;               \
    } while (0)

#define OFQ_REQUIRE(cond, ...)           \
    do {                                 \
        if (!(cond)) {                   \
            ofq_set_error(__VA_ARGS__);  \
            return OFQ_ERR_ARG;          \
        }                                \
    } while (0)
