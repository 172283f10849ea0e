// KD losses of the OFQ training recipe (reference src/quantization/utils.py:44-77: KLLossSoft, KDLossSoftandHard) with their
// gradients in ONE pass: per sample the hard cross entropy of the class-head logits, the soft cross entropy
// -sum softmax(t / T) log_softmax(z / T) of the distillation-head logits against the teacher's logits, and d loss / d logits
// for the mean over the batch. One warp per sample; the three rows of K logits are read once.
#include <cmath>
#include <cstdint>
#include "host_util.h"
#include "ofq_b200.h"

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// row statistics of z / T: max and log(sum exp)
__device__ __forceinline__ void row_lse(const float* __restrict__ z, int K, float invT, int lane, float* mx, float* lse) {
    float m = -INFINITY;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, __ldg(z + k) * invT);
    m = warp_max(m);
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += expf(__ldg(z + k) * invT - m);
    s = warp_sum(s);
    *mx = m;
    *lse = logf(s);
}

__global__ void __launch_bounds__(256)
kd_loss_kernel(const float* __restrict__ z_hard, const float* __restrict__ z_soft, const float* __restrict__ teacher,
               const long long* __restrict__ target, int B, int K, float invT, float inv_count, float* __restrict__ row_loss,
               float* __restrict__ dz_hard, float* __restrict__ dz_soft) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    const bool same = z_hard == z_soft;            // single-output student: both terms on the same logits
    float loss = 0.f;
    float mh = 0.f, lh = 0.f;
    long long y = -1;
    if (z_hard) {                                  // hard term: nn.CrossEntropyLoss (class indices), T does not apply
        const float* zh = z_hard + (long long)b * K;
        row_lse(zh, K, 1.0f, lane, &mh, &lh);
        y = target[b];
        loss += -(__ldg(zh + y) - mh - lh);
    }
    float ms = 0.f, ls = 0.f, mt = 0.f, lt = 0.f;
    if (teacher) {                                 // soft term
        const float* zs = z_soft + (long long)b * K;
        const float* tt = teacher + (long long)b * K;
        row_lse(zs, K, invT, lane, &ms, &ls);
        row_lse(tt, K, invT, lane, &mt, &lt);
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float p = expf(__ldg(tt + k) * invT - mt - lt);
            acc += p * (__ldg(zs + k) * invT - ms - ls);
        }
        loss -= warp_sum(acc);
    }
    if (lane == 0) row_loss[b] = loss;
    // gradients of mean_b(loss_b)
    for (int k = lane; k < K; k += 32) {
        float gh = 0.f, gs = 0.f;
        if (z_hard) gh = (expf(__ldg(z_hard + (long long)b * K + k) - mh - lh) - (k == y ? 1.f : 0.f)) * inv_count;
        if (teacher)
            gs = (expf(__ldg(z_soft + (long long)b * K + k) * invT - ms - ls) - expf(__ldg(teacher + (long long)b * K + k) * invT - mt - lt)) *
                 (invT * inv_count);
        if (same) {
            dz_hard[(long long)b * K + k] = gh + gs;
        } else {
            if (z_hard && dz_hard) dz_hard[(long long)b * K + k] = gh;
            if (teacher && dz_soft) dz_soft[(long long)b * K + k] = gs;
        }
    }
}

// deterministic mean of the per-sample losses (one CTA, fixed order)
__global__ void __launch_bounds__(256) kd_loss_mean_kernel(const float* __restrict__ row_loss, int B, float inv_count, float* __restrict__ out) {
    __shared__ float part[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) a += row_loss[i];
    part[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 256; ++i) s += part[i];
        *out = s * inv_count;
    }
}

}  // namespace

extern "C" int ofq_kd_loss(const float* z_hard, const float* z_soft, const float* teacher, const long long* target, int B, int K,
                           float T, float* row_loss, float* loss, float* dz_hard, float* dz_soft, void* stream) {
    OFQ_REQUIRE(B > 0 && K > 0 && T > 0.f && row_loss && loss, "ofq_kd_loss: bad argument");
    OFQ_REQUIRE(z_hard || teacher, "ofq_kd_loss: neither a hard nor a soft term requested");
    OFQ_REQUIRE(!z_hard || target, "ofq_kd_loss: the hard term needs class indices");
    OFQ_REQUIRE(!teacher || z_soft, "ofq_kd_loss: the soft term needs student logits");
    OFQ_REQUIRE(!(z_hard && teacher && z_hard == z_soft) || dz_hard, "ofq_kd_loss: shared logits need dz_hard");
    OFQ_CHECK_ARCH();
    cudaStream_t st = (cudaStream_t)stream;
    kd_loss_kernel<<<(B + 7) / 8, 256, 0, st>>>(z_hard, z_soft, teacher, target, B, K, 1.0f / T, 1.0f / (float)B, row_loss, dz_hard, dz_soft);
    kd_loss_mean_kernel<<<1, 256, 0, st>>>(row_loss, B, 1.0f / (float)B, loss);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}
