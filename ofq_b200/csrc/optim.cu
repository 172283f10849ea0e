// K7: confidence-guided-annealing freeze mask (cga.py:450-469) fused into the AdamW update
// (cga.py:953-1013 around torch.optim.AdamW). One pre-pass computes the per-row StatsQ scale and the global
// min/max rounding level; one pass then reads p, g, m, v and writes p, m, v (32 B/param in total).
#include "host_util.h"
#include "ofq_b200.h"
#include <climits>
#include <cmath>
#include <cstdint>

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// statsq.py:137-147 restated: b4_round = clamp(w/sf, -1, 1-1e-6) * n - 0.5
__device__ __forceinline__ float statsq_b4(float w, float sf, float n_levels) {
    float v = __fdiv_rn(w, sf);
    v = fminf(fmaxf(v, -1.0f), __fsub_rn(1.0f, 1e-6f));
    return __fsub_rn(__fmul_rn(v, n_levels), 0.5f);
}

__global__ void cga_init_kernel(int* kminmax) {
    kminmax[0] = INT_MAX;
    kminmax[1] = INT_MIN;
}

__global__ void __launch_bounds__(256)
cga_rowstat_kernel(const float* __restrict__ w, int rows, int cols, float n_levels, float* __restrict__ rowstat,
                   int* __restrict__ kminmax) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float* wr = w + (long long)row * cols;
    double acc = 0.0;
    for (int c = lane; c < cols; c += 32) acc += (double)fabsf(__ldg(wr + c));
    acc = warp_sum_d(acc);
    const float sf = __fmul_rn(2.0f, __fdiv_rn((float)acc, (float)cols));
    int kmin = INT_MAX, kmax = INT_MIN;
    for (int c = lane; c < cols; c += 32) {
        const int k = (int)rintf(statsq_b4(__ldg(wr + c), sf, n_levels));
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) {
        rowstat[row] = sf;
        atomicMin(kminmax, kmin);
        atomicMax(kminmax + 1, kmax);
    }
}

// cga.py:465-469: trainable iff for some integer i in [kmin, kmax): 0.5-BR <= b4 - i <= 0.5+BR (fp32 compares).
// Only i = floor(b4) can satisfy it.
__device__ __forceinline__ bool cga_frozen(float w, float sf, float n_levels, int kmin, int kmax, float lo_thr,
                                           float hi_thr) {
    const float b4 = statsq_b4(w, sf, n_levels);
    const float fi = floorf(b4);
    const int i = (int)fi;
    bool trainable = false;
    if (i >= kmin && i < kmax) {
        const float d = __fsub_rn(b4, fi);
        trainable = (d <= hi_thr) && (d >= lo_thr);
    }
    return !trainable;
}

__global__ void __launch_bounds__(256)
cga_mask_kernel(const float* __restrict__ w, long long numel, int cols, float n_levels,
                const float* __restrict__ rowstat, const int* __restrict__ kminmax, float lo_thr, float hi_thr,
                uint8_t* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numel) return;
    const int kmin = kminmax[0], kmax = kminmax[1];
    mask[i] = cga_frozen(__ldg(w + i), __ldg(rowstat + i / cols), n_levels, kmin, kmax, lo_thr, hi_thr) ? 1 : 0;
}

struct AdamScalars {
    float decay;        // 1 - lr * wd
    float one_m_b1;     // 1 - beta1
    float beta2;
    float one_m_b2;
    float step_size;    // lr / (1 - beta1^t)
    float bc2_sqrt;     // sqrt(1 - beta2^t)
    float eps;
    double lr, beta1_d, beta2_d;   // for the device-side step counter (CUDA-graph replay)
};

__global__ void counter_increment_kernel(int* c) { *c += 1; }

// torch/optim/adamw.py::_single_tensor_adamw element-wise math.
__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamScalars& a, bool frozen) {
    if (frozen) g = 0.f;                                    // cga.py:958 grad * freeze_idx * 0
    const float p_new0 = __fmul_rn(p, a.decay);             // param.mul_(1 - lr * wd): a rounded product of its own (never
                                                            // contracted into the update below, in any of the kernels)
    m = fmaf(a.one_m_b1, g - m, m);                         // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.one_m_b2 * g, g, v * a.beta2);               // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    const float p_new = __fmaf_rn(-a.step_size, __fdiv_rn(m, denom), p_new0);   // param.addcdiv_(exp_avg, denom, value=-step_size)
    if (!frozen) p = p_new;                                 // cga.py:1002-1005 restores frozen weights bit-for-bit
}

template <bool MASKED>
__global__ void __launch_bounds__(256)
cga_adamw_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, long long numel, int cols, float n_levels,
                 const float* __restrict__ rowstat, const int* __restrict__ kminmax, float lo_thr, float hi_thr,
                 AdamScalars a, const int* __restrict__ step_dev, uint8_t* __restrict__ mask_out) {
    if (step_dev) {   // bias corrections from a device-resident step count: the launch can be replayed by a CUDA graph
        __shared__ float bc[2];
        if (threadIdx.x == 0) {
            const double t = (double)(*step_dev);
            bc[0] = (float)(a.lr / (1.0 - pow(a.beta1_d, t)));
            bc[1] = (float)sqrt(1.0 - pow(a.beta2_d, t));
        }
        __syncthreads();
        a.step_size = bc[0];
        a.bc2_sqrt = bc[1];
    }
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= numel) return;
    int kmin = 0, kmax = 0;
    if (MASKED) { kmin = kminmax[0]; kmax = kminmax[1]; }
    if (i4 + 4 <= numel && (cols % 4 == 0 || !MASKED)) {
        float4 pp = *reinterpret_cast<float4*>(p + i4);
        const float4 gg = __ldg(reinterpret_cast<const float4*>(grad + i4));
        float4 mm = *reinterpret_cast<float4*>(m + i4);
        float4 vv = *reinterpret_cast<float4*>(v + i4);
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
        const float sf = MASKED ? __ldg(rowstat + i4 / cols) : 1.f;
        uint8_t fr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool frozen = MASKED ? cga_frozen(pa[e], sf, n_levels, kmin, kmax, lo_thr, hi_thr) : false;
            fr[e] = frozen;
            adamw_elem(pa[e], ga[e], ma[e], va[e], a, frozen);
        }
        *reinterpret_cast<float4*>(p + i4) = pp;
        *reinterpret_cast<float4*>(m + i4) = mm;
        *reinterpret_cast<float4*>(v + i4) = vv;
        if (MASKED && mask_out)
            *reinterpret_cast<uint32_t*>(mask_out + i4) = fr[0] | (fr[1] << 8) | (fr[2] << 16) | ((uint32_t)fr[3] << 24);
    } else {
        for (long long i = i4; i < numel && i < i4 + 4; ++i) {
            const bool frozen = MASKED ? cga_frozen(p[i], __ldg(rowstat + i / cols), n_levels, kmin, kmax, lo_thr, hi_thr) : false;
            float pv = p[i], mv = m[i], vv = v[i];
            adamw_elem(pv, grad[i], mv, vv, a, frozen);
            p[i] = pv; m[i] = mv; v[i] = vv;
            if (MASKED && mask_out) mask_out[i] = frozen;
        }
    }
}

// Multi-tensor plain AdamW: one launch walks a table of (pointer, length, weight-decay) entries.
struct AdamWEntry {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long numel;
    float wd;          // weight decay of this parameter's group (the decay factor 1 - lr * wd is formed in the kernel: the
                       // table does not change with the learning-rate schedule)
    int first_block;   // first CTA that works on this entry
};

// Multi-tensor CGA: every freeze-masked weight of the model in ONE pre-pass launch and ONE update launch.
struct CgaEntry {
    float* p;
    const float* g;
    float* m;
    float* v;
    float* rowstat;    // [rows] scratch: per-row StatsQ scale
    int* kminmax;      // [2] scratch: global min / max rounding level of this weight
    int rows, cols;
    float wd;
    int first_block;   // update kernel: running sum of ceil(rows * cols / 1024)
    int first_rowblock;// pre-pass kernel: running sum of ceil(rows / 8)
    int pad;
};

template <typename E>
__device__ __forceinline__ int find_entry(const E* table, int n_entries, int block, int E::*first) {
    int lo = 0, hi = n_entries - 1;                 // last entry whose first block <= block
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].*first <= block) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void cga_init_multi_kernel(const CgaEntry* __restrict__ table, int n_entries) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_entries) { table[i].kminmax[0] = INT_MAX; table[i].kminmax[1] = INT_MIN; }
}

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamWEntry* __restrict__ table, int n_entries, AdamScalars a, const int* __restrict__ step_dev) {
    __shared__ float bc[2];
    __shared__ int entry_s;
    if (threadIdx.x == 0) {
        if (step_dev) {
            const double t = (double)(*step_dev);
            bc[0] = (float)(a.lr / (1.0 - pow(a.beta1_d, t)));
            bc[1] = (float)sqrt(1.0 - pow(a.beta2_d, t));
        } else {
            bc[0] = a.step_size;
            bc[1] = a.bc2_sqrt;
        }
        int lo = 0, hi = n_entries - 1;             // last entry whose first_block <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        entry_s = lo;
    }
    __syncthreads();
    a.step_size = bc[0];
    a.bc2_sqrt = bc[1];
    const AdamWEntry e = table[entry_s];
    a.decay = (float)(1.0 - a.lr * (double)e.wd);
    const long long base = ((long long)(blockIdx.x - e.first_block) * blockDim.x + threadIdx.x) * 4;
    if (base >= e.numel) return;
    if (base + 4 <= e.numel && (((uintptr_t)e.p | (uintptr_t)e.g | (uintptr_t)e.m | (uintptr_t)e.v) & 15) == 0) {
        float4 pp = *reinterpret_cast<float4*>(e.p + base);
        const float4 gg = __ldg(reinterpret_cast<const float4*>(e.g + base));
        float4 mm = *reinterpret_cast<float4*>(e.m + base);
        float4 vv = *reinterpret_cast<float4*>(e.v + base);
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) adamw_elem(pa[k], ga[k], ma[k], va[k], a, false);
        *reinterpret_cast<float4*>(e.p + base) = pp;
        *reinterpret_cast<float4*>(e.m + base) = mm;
        *reinterpret_cast<float4*>(e.v + base) = vv;
    } else {
        for (long long i = base; i < e.numel && i < base + 4; ++i) {
            float pv = e.p[i], mv = e.m[i], vv = e.v[i];
            adamw_elem(pv, e.g[i], mv, vv, a, false);
            e.p[i] = pv; e.m[i] = mv; e.v[i] = vv;
        }
    }
}

// pre-pass of all masked weights: block = 8 rows of one entry (cga_rowstat_kernel's work per row)
__global__ void __launch_bounds__(256)
cga_rowstat_multi_kernel(const CgaEntry* __restrict__ table, int n_entries, float n_levels) {
    __shared__ int entry_s;
    if (threadIdx.x == 0) entry_s = find_entry(table, n_entries, (int)blockIdx.x, &CgaEntry::first_rowblock);
    __syncthreads();
    const CgaEntry e = table[entry_s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = ((int)blockIdx.x - e.first_rowblock) * 8 + warp;
    if (row >= e.rows) return;
    const float* wr = e.p + (long long)row * e.cols;
    double acc = 0.0;
    for (int c = lane; c < e.cols; c += 32) acc += (double)fabsf(__ldg(wr + c));
    acc = warp_sum_d(acc);
    const float sf = __fmul_rn(2.0f, __fdiv_rn((float)acc, (float)e.cols));
    int kmin = INT_MAX, kmax = INT_MIN;
    for (int c = lane; c < e.cols; c += 32) {
        const int k = (int)rintf(statsq_b4(__ldg(wr + c), sf, n_levels));
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) {
        e.rowstat[row] = sf;
        atomicMin(e.kminmax, kmin);
        atomicMax(e.kminmax + 1, kmax);
    }
}

__global__ void __launch_bounds__(256)
cga_adamw_multi_kernel(const CgaEntry* __restrict__ table, int n_entries, float n_levels, float lo_thr, float hi_thr,
                       AdamScalars a, const int* __restrict__ step_dev) {
    __shared__ float bc[2];
    __shared__ int entry_s;
    if (threadIdx.x == 0) {
        if (step_dev) {
            const double t = (double)(*step_dev);
            bc[0] = (float)(a.lr / (1.0 - pow(a.beta1_d, t)));
            bc[1] = (float)sqrt(1.0 - pow(a.beta2_d, t));
        } else {
            bc[0] = a.step_size;
            bc[1] = a.bc2_sqrt;
        }
        entry_s = find_entry(table, n_entries, (int)blockIdx.x, &CgaEntry::first_block);
    }
    __syncthreads();
    a.step_size = bc[0];
    a.bc2_sqrt = bc[1];
    const CgaEntry e = table[entry_s];
    a.decay = (float)(1.0 - a.lr * (double)e.wd);
    const long long numel = (long long)e.rows * e.cols;
    const long long base = ((long long)((int)blockIdx.x - e.first_block) * blockDim.x + threadIdx.x) * 4;
    if (base >= numel) return;
    const int kmin = e.kminmax[0], kmax = e.kminmax[1];
    if (base + 4 <= numel && e.cols % 4 == 0) {
        float4 pp = *reinterpret_cast<float4*>(e.p + base);
        const float4 gg = __ldg(reinterpret_cast<const float4*>(e.g + base));
        float4 mm = *reinterpret_cast<float4*>(e.m + base);
        float4 vv = *reinterpret_cast<float4*>(e.v + base);
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
        const float sf = __ldg(e.rowstat + base / e.cols);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            adamw_elem(pa[k], ga[k], ma[k], va[k], a, cga_frozen(pa[k], sf, n_levels, kmin, kmax, lo_thr, hi_thr));
        *reinterpret_cast<float4*>(e.p + base) = pp;
        *reinterpret_cast<float4*>(e.m + base) = mm;
        *reinterpret_cast<float4*>(e.v + base) = vv;
    } else {
        for (long long i = base; i < numel && i < base + 4; ++i) {
            float pv = e.p[i], mv = e.m[i], vv = e.v[i];
            adamw_elem(pv, e.g[i], mv, vv, a, cga_frozen(pv, __ldg(e.rowstat + i / e.cols), n_levels, kmin, kmax, lo_thr, hi_thr));
            e.p[i] = pv; e.m[i] = mv; e.v[i] = vv;
        }
    }
}

}  // namespace

static int cga_prepass(const float* w, int rows, int cols, int bits, float* rowstat, int* kminmax, cudaStream_t st) {
    cga_init_kernel<<<1, 1, 0, st>>>(kminmax);
    cga_rowstat_kernel<<<(rows + 7) / 8, 256, 0, st>>>(w, rows, cols, (float)(1 << (bits - 1)), rowstat, kminmax);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_cga_mask(const float* w, int rows, int cols, int bits, double boundary_range, uint8_t* mask,
                            float* rowstat, int* kminmax, void* stream) {
    OFQ_REQUIRE(w && mask && rowstat && kminmax && rows > 0 && cols > 0 && bits >= 2 && bits <= 7, "ofq_cga_mask: bad argument");
    OFQ_CHECK_ARCH();
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cga_prepass(w, rows, cols, bits, rowstat, kminmax, st);
    if (rc) return rc;
    const long long numel = (long long)rows * cols;
    cga_mask_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, st>>>(
        w, numel, cols, (float)(1 << (bits - 1)), rowstat, kminmax, (float)(0.5 - boundary_range),
        (float)(0.5 + boundary_range), mask);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_cga_adamw(float* p, const float* grad, float* exp_avg, float* exp_avg_sq, long long numel,
                             int rows, int cols, int step, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int bits, double boundary_range, float* rowstat, int* kminmax,
                             uint8_t* mask_out, const int* step_dev, void* stream) {
    OFQ_REQUIRE(p && grad && exp_avg && exp_avg_sq && numel > 0 && (step >= 1 || step_dev), "ofq_cga_adamw: bad argument");
    if (step < 1) step = 1;
    OFQ_REQUIRE((uintptr_t)p % 16 == 0 && (uintptr_t)grad % 16 == 0 && (uintptr_t)exp_avg % 16 == 0 &&
                (uintptr_t)exp_avg_sq % 16 == 0, "ofq_cga_adamw: tensors must be 16-byte aligned");
    OFQ_CHECK_ARCH();
    cudaStream_t st = (cudaStream_t)stream;
    AdamScalars a;
    a.decay = (float)(1.0 - lr * weight_decay);
    a.one_m_b1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.one_m_b2 = (float)(1.0 - beta2);
    a.step_size = (float)(lr / (1.0 - std::pow(beta1, (double)step)));
    a.bc2_sqrt = (float)std::sqrt(1.0 - std::pow(beta2, (double)step));
    a.eps = (float)eps;
    a.lr = lr; a.beta1_d = beta1; a.beta2_d = beta2;
    const unsigned grid = (unsigned)((numel + 1023) / 1024);
    if (bits > 0) {
        OFQ_REQUIRE(rows > 0 && cols > 0 && (long long)rows * cols == numel && rowstat && kminmax && bits >= 2 && bits <= 7,
                    "ofq_cga_adamw: masked update needs a 2-D weight (rows*cols == numel), rowstat and kminmax scratch");
        OFQ_REQUIRE(!mask_out || (uintptr_t)mask_out % 4 == 0, "ofq_cga_adamw: mask_out alignment");
        int rc = cga_prepass(p, rows, cols, bits, rowstat, kminmax, st);
        if (rc) return rc;
        cga_adamw_kernel<true><<<grid, 256, 0, st>>>(p, grad, exp_avg, exp_avg_sq, numel, cols, (float)(1 << (bits - 1)),
                                                     rowstat, kminmax, (float)(0.5 - boundary_range),
                                                     (float)(0.5 + boundary_range), a, step_dev, mask_out);
    } else {
        cga_adamw_kernel<false><<<grid, 256, 0, st>>>(p, grad, exp_avg, exp_avg_sq, numel, 1, 1.f, nullptr, nullptr, 0.f,
                                                      0.f, a, step_dev, nullptr);
    }
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_counter_increment(int* counter, void* stream) {
    OFQ_REQUIRE(counter, "ofq_counter_increment: null pointer");
    OFQ_CHECK_ARCH();
    counter_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

// table: device array of n_entries records {p, g, m, v (pointers), numel (int64), decay (float), first_block (int)}
// laid out as the AdamWEntry struct above (48 bytes); total_blocks = sum of ceil(numel / 1024).
extern "C" int ofq_adamw_multi(const void* table, int n_entries, int total_blocks, int step, double lr, double beta1,
                               double beta2, double eps, const int* step_dev, void* stream) {
    OFQ_REQUIRE(table && n_entries > 0 && total_blocks > 0 && (step >= 1 || step_dev), "ofq_adamw_multi: bad argument");
    static_assert(sizeof(AdamWEntry) == 48, "AdamWEntry layout is part of the C-ABI");
    OFQ_CHECK_ARCH();
    if (step < 1) step = 1;
    AdamScalars a;
    a.decay = 1.f;
    a.one_m_b1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.one_m_b2 = (float)(1.0 - beta2);
    a.step_size = (float)(lr / (1.0 - std::pow(beta1, (double)step)));
    a.bc2_sqrt = (float)std::sqrt(1.0 - std::pow(beta2, (double)step));
    a.eps = (float)eps;
    a.lr = lr; a.beta1_d = beta1; a.beta2_d = beta2;
    adamw_multi_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>((const AdamWEntry*)table, n_entries, a, step_dev);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

// table: device array of n_entries 72-byte CgaEntry records
//   { float* p; const float* g; float* m; float* v; float* rowstat; int32* kminmax; int32 rows, cols; float wd;
//     int32 first_block (running sum of ceil(rows*cols / 1024)); int32 first_rowblock (running sum of ceil(rows / 8)); int32 pad }
// Three launches for ALL freeze-masked weights of a model (scratch init, per-row statistics + level range, masked update)
// instead of three per weight.
extern "C" int ofq_cga_adamw_multi(const void* table, int n_entries, int total_blocks, int total_rowblocks, int step, double lr,
                                   double beta1, double beta2, double eps, int bits, double boundary_range, const int* step_dev,
                                   void* stream) {
    OFQ_REQUIRE(table && n_entries > 0 && total_blocks > 0 && total_rowblocks > 0 && (step >= 1 || step_dev) && bits >= 2 && bits <= 7,
                "ofq_cga_adamw_multi: bad argument");
    static_assert(sizeof(CgaEntry) == 72, "CgaEntry layout is part of the C-ABI");
    OFQ_CHECK_ARCH();
    if (step < 1) step = 1;
    cudaStream_t st = (cudaStream_t)stream;
    AdamScalars a;
    a.decay = 1.f;
    a.one_m_b1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.one_m_b2 = (float)(1.0 - beta2);
    a.step_size = (float)(lr / (1.0 - std::pow(beta1, (double)step)));
    a.bc2_sqrt = (float)std::sqrt(1.0 - std::pow(beta2, (double)step));
    a.eps = (float)eps;
    a.lr = lr; a.beta1_d = beta1; a.beta2_d = beta2;
    const CgaEntry* t = (const CgaEntry*)table;
    const float n_levels = (float)(1 << (bits - 1));
    cga_init_multi_kernel<<<(n_entries + 127) / 128, 128, 0, st>>>(t, n_entries);
    cga_rowstat_multi_kernel<<<total_rowblocks, 256, 0, st>>>(t, n_entries, n_levels);
    cga_adamw_multi_kernel<<<total_blocks, 256, 0, st>>>(t, n_entries, n_levels, (float)(0.5 - boundary_range), (float)(0.5 + boundary_range),
                                                         a, step_dev);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}
