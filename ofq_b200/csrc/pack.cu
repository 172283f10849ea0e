// Deployment export (no reference counterpart - the reference stops at fake-quant floats): the bit-exact
// StatsQ weight codes packed to their true width. A b-bit code 2k+1 (k in [-n, n-1], n = 2^(b-1)) is stored as u = k + n in b bits,
// eight codes per b bytes (little-endian bit order), rows padded to a multiple of eight codes.
#include <cstdint>
#include "host_util.h"
#include "ofq_b200.h"

namespace {

__global__ void __launch_bounds__(256)
pack_codes_kernel(const int8_t* __restrict__ codes, long long rows, int cols, long long ld, int bits, int n, uint8_t* __restrict__ out,
                  long long ld_out) {
    const int groups = (cols + 7) / 8;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * groups) return;
    const long long r = i / groups;
    const int g = (int)(i - r * groups);
    unsigned long long word = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = g * 8 + e;
        unsigned u = 0;
        if (c < cols) u = (unsigned)(((int)codes[r * ld + c] - 1) / 2 + n) & ((1u << bits) - 1u);      // code = 2k+1 -> k + n
        word |= (unsigned long long)u << (e * bits);
    }
    uint8_t* o = out + r * ld_out + (long long)g * bits;
    for (int b = 0; b < bits; ++b) o[b] = (uint8_t)(word >> (8 * b));
}

__global__ void __launch_bounds__(256)
unpack_codes_kernel(const uint8_t* __restrict__ in, long long ld_in, long long rows, int cols, int bits, int n, int8_t* __restrict__ codes,
                    long long ld) {
    const int groups = (cols + 7) / 8;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * groups) return;
    const long long r = i / groups;
    const int g = (int)(i - r * groups);
    const uint8_t* p = in + r * ld_in + (long long)g * bits;
    unsigned long long word = 0;
    for (int b = 0; b < bits; ++b) word |= (unsigned long long)p[b] << (8 * b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = g * 8 + e;
        if (c < cols) {
            const int u = (int)((word >> (e * bits)) & ((1ull << bits) - 1ull));
            codes[r * ld + c] = (int8_t)(2 * (u - n) + 1);
        }
    }
}

}  // namespace

extern "C" long long ofq_packed_row_bytes(int cols, int bits) { return (long long)((cols + 7) / 8) * bits; }

extern "C" int ofq_pack_codes(const int8_t* codes, long long rows, int cols, long long ld, int bits, uint8_t* out, long long ld_out,
                              void* stream) {
    OFQ_REQUIRE(codes && out && rows > 0 && cols > 0 && bits >= 2 && bits <= 7 && ld >= cols && ld_out >= ofq_packed_row_bytes(cols, bits),
                "ofq_pack_codes: bad argument");
    OFQ_CHECK_ARCH();
    const long long total = rows * ((cols + 7) / 8);
    pack_codes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(codes, rows, cols, ld, bits, 1 << (bits - 1), out, ld_out);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_unpack_codes(const uint8_t* in, long long ld_in, long long rows, int cols, int bits, int8_t* codes, long long ld,
                                void* stream) {
    OFQ_REQUIRE(codes && in && rows > 0 && cols > 0 && bits >= 2 && bits <= 7 && ld >= cols && ld_in >= ofq_packed_row_bytes(cols, bits),
                "ofq_unpack_codes: bad argument");
    OFQ_CHECK_ARCH();
    const long long total = rows * ((cols + 7) / 8);
    unpack_codes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, rows, cols, bits, 1 << (bits - 1), codes, ld);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}
