// Host-glue kernels around the quantized layers: LayerNorm forward / backward in fp32.
// torch's LayerNorm backward is ~5x off the HBM roofline for 384-wide rows; these keep the residual stream in fp32 and
// reproduce nn.LayerNorm (biased variance, eps inside the sqrt) to fp32 round-off.
#include "host_util.h"
#include "ofq_b200.h"
#include <cstdint>

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row, two passes over the row (second pass hits L1)
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, long long rows, int cols, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
    const int c4 = cols >> 2;
    float s = 0.f;
    for (int i = lane; i < c4; i += 32) {
        const float4 v = xr[i];
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = wsum(s) / (float)cols;
    float q = 0.f;
    for (int i = lane; i < c4; i += 32) {
        const float4 v = xr[i];
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(wsum(q) / (float)cols + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * cols);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    for (int i = lane; i < c4; i += 32) {
        const float4 v = xr[i], g = __ldg(g4 + i), b = __ldg(b4 + i);
        yr[i] = make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y,
                            (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w);
    }
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

// x_new = x + add (the residual add in front of a pre-norm LayerNorm, deit_vision_transformer.py:156-163) and
// y = LayerNorm(x_new) in ONE pass: rows of at most 512 columns stay in registers (4 x 16 B per lane), so the sum is read
// once and never re-read. Same summation order as layernorm_fwd_kernel: identical statistics and outputs.
__global__ void __launch_bounds__(256)
layernorm_fwd_add_kernel(const float* __restrict__ x, const float* __restrict__ add, long long rows, int cols,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         float* __restrict__ xsum, float* __restrict__ y, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
    const float4* ar = reinterpret_cast<const float4*>(add + row * cols);
    float4* sr = reinterpret_cast<float4*>(xsum + row * cols);
    const int c4 = cols >> 2;
    float4 v[4];
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = lane + 32 * p;
        v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < c4) {
            const float4 a = __ldg(xr + i), b = __ldg(ar + i);
            v[p] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            sr[i] = v[p];
            s += (v[p].x + v[p].y) + (v[p].z + v[p].w);
        }
    }
    const float mean = wsum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (lane + 32 * p < c4) {
            const float a = v[p].x - mean, b = v[p].y - mean, c = v[p].z - mean, d = v[p].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(wsum(q) / (float)cols + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * cols);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = lane + 32 * p;
        if (i < c4) {
            const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
            yr[i] = make_float4((v[p].x - mean) * rstd * g.x + b.x, (v[p].y - mean) * rstd * g.y + b.y,
                                (v[p].z - mean) * rstd * g.z + b.z, (v[p].w - mean) * rstd * g.w + b.w);
        }
    }
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

constexpr int kChunk = 512;
constexpr int kMaxBlocks = 4 * 148;

// grid (row blocks, column chunks of 512). dx for the chunk, per-block partial sums of dgamma / dbeta.
template <bool ONE_CHUNK>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, long long rows, int cols,
                     const float* __restrict__ res, float* __restrict__ dx, float* __restrict__ part,
                     float* __restrict__ blockmax) {
    __shared__ float col_s[2][kChunk];
    __shared__ float bmax_s[8];
    float tmax = 0.f;             // max |dx| written by this thread: the fp16 range scale of the next GEMM operand needs no pass
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rpb = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rpb, r1 = min(rows, r0 + rpb);
    const int cbase = blockIdx.y * kChunk;
    for (int i = threadIdx.x; i < 2 * kChunk; i += blockDim.x) (&col_s[0][0])[i] = 0.f;
    __syncthreads();
    float a_g[4][4], a_b[4][4];
    float4 gv[4];
    bool ok[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int col = cbase + p * 128 + lane * 4;
        ok[p] = col < cols;
        gv[p] = ok[p] ? __ldg(reinterpret_cast<const float4*>(gamma + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) a_g[p][e] = a_b[p][e] = 0.f;
    }
    const int c4 = cols >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    if (ONE_CHUNK) {
        // the whole row fits the 512-column chunk (C = 384): ONE pass, the row stays in registers (8 x 16 B in flight per lane)
        for (long long r = r0 + warp; r < r1; r += 8) {
            float4 d[4], v[4];
            const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
            if (res && (lane & 7) == 0) {       // the residual gradient is consumed after the row reductions: start it now
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (ok[p]) asm volatile("prefetch.global.L2 [%0];" ::"l"(res + r * cols + (p * 32 + lane) * 4));
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                d[p] = v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[p]) {
                    d[p] = __ldg(reinterpret_cast<const float4*>(dy + r * cols) + p * 32 + lane);
                    v[p] = __ldg(reinterpret_cast<const float4*>(x + r * cols) + p * 32 + lane);
                }
            }
            float c1 = 0.f, c2 = 0.f;               // sum(dy*g), sum(dy*g*(x-mu)) over the whole row
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float t0 = d[p].x * gv[p].x, t1 = d[p].y * gv[p].y, t2 = d[p].z * gv[p].z, t3 = d[p].w * gv[p].w;
                c1 += (t0 + t1) + (t2 + t3);
                c2 += ok[p] ? (t0 * (v[p].x - mu) + t1 * (v[p].y - mu)) + (t2 * (v[p].z - mu) + t3 * (v[p].w - mu)) : 0.f;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                c2 += __shfl_xor_sync(0xffffffffu, c2, o);
            }
            c1 = c1 / (float)cols;
            c2 = c2 * rs / (float)cols;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (!ok[p]) continue;
                const float dd[4] = {d[p].x, d[p].y, d[p].z, d[p].w};
                const float vv[4] = {v[p].x, v[p].y, v[p].z, v[p].w};
                const float gg[4] = {gv[p].x, gv[p].y, gv[p].z, gv[p].w};
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xh = (vv[e] - mu) * rs;
                    o[e] = rs * (dd[e] * gg[e] - c1 - xh * c2);
                    a_g[p][e] += dd[e] * xh;
                    a_b[p][e] += dd[e];
                }
                if (res) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(res + r * cols) + p * 32 + lane);
                    o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
                }
                tmax = fmaxf(fmaxf(tmax, fmaxf(fabsf(o[0]), fabsf(o[1]))), fmaxf(fabsf(o[2]), fabsf(o[3])));
                reinterpret_cast<float4*>(dx + r * cols)[p * 32 + lane] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    } else
    for (long long row = r0 + warp; row < r1; row += 8) {
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        const float4* dyr = reinterpret_cast<const float4*>(dy + row * cols);
        const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
        float c1 = 0.f, c2 = 0.f;                       // sum(dy*g), sum(dy*g*xhat) over the whole row
        for (int i = lane; i < c4; i += 32) {
            const float4 d = __ldg(dyr + i), v = __ldg(xr + i), g = __ldg(g4 + i);
            const float t0 = d.x * g.x, t1 = d.y * g.y, t2 = d.z * g.z, t3 = d.w * g.w;
            c1 += (t0 + t1) + (t2 + t3);
            c2 += (t0 * (v.x - mu) + t1 * (v.y - mu)) + (t2 * (v.z - mu) + t3 * (v.w - mu));
        }
        c1 = wsum(c1) / (float)cols;
        c2 = wsum(c2) * rs / (float)cols;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (!ok[p]) continue;
            const int i = (cbase >> 2) + p * 32 + lane;
            const float4 d = __ldg(dyr + i), v = __ldg(xr + i);
            const float dd[4] = {d.x, d.y, d.z, d.w}, vv[4] = {v.x, v.y, v.z, v.w};
            const float gg[4] = {gv[p].x, gv[p].y, gv[p].z, gv[p].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xh = (vv[e] - mu) * rs;
                o[e] = rs * (dd[e] * gg[e] - c1 - xh * c2);
                a_g[p][e] += dd[e] * xh;
                a_b[p][e] += dd[e];
            }
            if (res) {      // gradient arriving over the residual connection around this LayerNorm: summed here, not by a separate add
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(res + row * cols) + i);
                o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
            }
            reinterpret_cast<float4*>(dx + row * cols)[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
        if (ok[p])
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int si = (p * 4 + e) * 32 + lane;
                atomicAdd(&col_s[0][si], a_g[p][e]);
                atomicAdd(&col_s[1][si], a_b[p][e]);
            }
    __syncthreads();
    if (blockmax && ONE_CHUNK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if (lane == 0) bmax_s[warp] = tmax;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = bmax_s[0];
#pragma unroll
            for (int w2 = 1; w2 < 8; ++w2) m = fmaxf(m, bmax_s[w2]);
            blockmax[blockIdx.x] = m;
        }
    }
    float* pp = part + (long long)blockIdx.x * 2 * cols;
    for (int i = threadIdx.x; i < kChunk; i += blockDim.x)
        if (cbase + i < cols) {
            const int si = ((i >> 7) * 4 + (i & 3)) * 32 + ((i >> 2) & 31);
            pp[cbase + i] = col_s[0][si];
            pp[cols + cbase + i] = col_s[1][si];
        }
}

// block = 32 columns x 32 slices of the partial rows (a few hundred partial rows: ~20 dependent loads per thread)
__global__ void __launch_bounds__(1024)
colpart_reduce2_kernel(const float* __restrict__ part, int cols, long long nblk, float* __restrict__ out0,
                       float* __restrict__ out1) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx, vecid = blockIdx.y;
    float acc = 0.f;
    if (col < cols)
        for (long long b = ty; b < nblk; b += 32) acc += part[(b * 2 + vecid) * cols + col];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < cols) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += red[k][tx];
        (vecid == 0 ? out0 : out1)[col] = s;
    }
}

long long ln_nblk(long long rows) {
    const long long n = (rows + 31) / 32;
    return n < kMaxBlocks ? n : kMaxBlocks;
}

}  // namespace

extern "C" int ofq_layernorm_fwd(const float* x, long long rows, int cols, const float* gamma, const float* beta,
                                 float eps, float* y, float* mean, float* rstd, void* stream) {
    OFQ_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0 && cols > 0, "ofq_layernorm_fwd: bad argument");
    OFQ_REQUIRE(cols % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)gamma % 16 == 0 &&
                (uintptr_t)beta % 16 == 0, "ofq_layernorm_fwd: cols must be a multiple of 4 and pointers 16-byte aligned");
    OFQ_CHECK_ARCH();
    layernorm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, gamma, beta, eps, y, mean, rstd);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_layernorm_fwd_add(const float* x, const float* add, long long rows, int cols, const float* gamma,
                                     const float* beta, float eps, float* xsum, float* y, float* mean, float* rstd,
                                     void* stream) {
    OFQ_REQUIRE(x && add && gamma && beta && xsum && y && mean && rstd && rows > 0 && cols > 0, "ofq_layernorm_fwd_add: bad argument");
    OFQ_REQUIRE(cols % 4 == 0 && cols <= 512, "ofq_layernorm_fwd_add: rows of at most 512 columns, a multiple of 4");
    OFQ_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)add % 16 == 0 && (uintptr_t)xsum % 16 == 0 && (uintptr_t)y % 16 == 0 &&
                (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0, "ofq_layernorm_fwd_add: pointers must be 16-byte aligned");
    OFQ_CHECK_ARCH();
    layernorm_fwd_add_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, add, rows, cols, gamma, beta, eps,
                                                                                         xsum, y, mean, rstd);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" long long ofq_layernorm_bwd_workspace(long long rows, int cols) { return ln_nblk(rows) * 2 * cols; }

extern "C" int ofq_layernorm_bwd_res(const float* dy, const float* x, const float* gamma, const float* mean,
                                     const float* rstd, long long rows, int cols, const float* res, float* dx,
                                     float* dgamma, float* dbeta, float* workspace, void* stream) {
    return ofq_layernorm_bwd_max(dy, x, gamma, mean, rstd, rows, cols, res, dx, dgamma, dbeta, workspace, nullptr, stream);
}

extern "C" long long ofq_layernorm_bwd_nmax(long long rows, int cols) { return cols <= kChunk ? ln_nblk(rows) : 0; }

extern "C" int ofq_layernorm_bwd_max(const float* dy, const float* x, const float* gamma, const float* mean,
                                     const float* rstd, long long rows, int cols, const float* res, float* dx,
                                     float* dgamma, float* dbeta, float* workspace, float* blockmax, void* stream) {
    OFQ_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace && rows > 0 && cols > 0,
                "ofq_layernorm_bwd: bad argument");
    OFQ_REQUIRE(cols % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)dy % 16 == 0 && (uintptr_t)dx % 16 == 0 &&
                (uintptr_t)gamma % 16 == 0 && (uintptr_t)res % 16 == 0,
                "ofq_layernorm_bwd: cols must be a multiple of 4 and pointers 16-byte aligned");
    OFQ_CHECK_ARCH();
    cudaStream_t st = (cudaStream_t)stream;
    const long long nblk = ln_nblk(rows);
    dim3 grid((unsigned)nblk, (cols + kChunk - 1) / kChunk);
    OFQ_REQUIRE(!blockmax || grid.y == 1, "ofq_layernorm_bwd_max: block maxima are produced for rows of at most 512 columns");
    if (grid.y == 1) layernorm_bwd_kernel<true><<<grid, 256, 0, st>>>(dy, x, gamma, mean, rstd, rows, cols, res, dx, workspace, blockmax);
    else layernorm_bwd_kernel<false><<<grid, 256, 0, st>>>(dy, x, gamma, mean, rstd, rows, cols, res, dx, workspace, nullptr);
    dim3 g2((cols + 31) / 32, 2);
    colpart_reduce2_kernel<<<g2, 1024, 0, st>>>(workspace, cols, nblk, dgamma, dbeta);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, long long rows, int cols, float* dx, float* dgamma, float* dbeta,
                                 float* workspace, void* stream) {
    return ofq_layernorm_bwd_res(dy, x, gamma, mean, rstd, rows, cols, nullptr, dx, dgamma, dbeta, workspace, stream);
}
