// Error reporting, architecture gate and TMA tensor-map encoding for the ofq_b200 C-ABI.
#include "host_util.h"
#include <cudaTypedefs.h>
#include <mutex>

static thread_local char g_err[512] = "";

void ofq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ofq_last_error(void) { return g_err; }
extern "C" int ofq_version(void) { return 1; }

static int g_arch_dev = -1, g_arch_ok = 0, g_sms = 0;

int ofq_check_arch() {
    int dev = -1;
    OFQ_CUDA(cudaGetDevice(&dev));
    if (dev != g_arch_dev) {
        int major = 0, sms = 0;
        OFQ_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        OFQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        g_arch_ok = (major == 10);
        g_sms = sms;
        g_arch_dev = dev;
    }
    if (!g_arch_ok) {
        ofq_set_error("ofq_b200 kernels are built for sm_100a only; current device is not compute capability 10.x");
        return OFQ_ERR_ARCH;
    }
    return 0;
}

int ofq_num_sms() {
    if (g_arch_dev < 0) ofq_check_arch();
    return g_sms > 0 ? g_sms : 148;
}

extern "C" int ofq_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        ofq_set_error("no CUDA device");
        return OFQ_ERR_CUDA;
    }
    return ofq_check_arch() == 0 ? 1 : 0;
}

// libcuda is resolved at run time (the library must load on a machine without a driver so that the
// symbol-export test can run there).
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

int ofq_encode_tensor_map(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, void* addr,
                          const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                          const cuuint32_t* estr) {
    return ofq_encode_tensor_map_sw(tm, dtype, rank, addr, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
}

int ofq_encode_tensor_map_sw(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, void* addr,
                             const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                             const cuuint32_t* estr, CUtensorMapSwizzle swizzle) {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    if (!g_encode) {
        ofq_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return OFQ_ERR_CUDA;
    }
    CUresult r = g_encode(tm, dtype, rank, addr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ofq_set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu,%llu strides %llu,%llu,%llu,%llu box %u,%u)",
                      (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                      (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)dims[4],
                      (unsigned long long)strides[0], (unsigned long long)strides[1],
                      (unsigned long long)strides[2], (unsigned long long)strides[3], box[0], box[1]);
        return OFQ_ERR_CUDA;
    }
    return 0;
}
