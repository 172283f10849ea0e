// Softmax + unsigned LSQ quantization of attention probabilities, forward and backward
// (reference attention.py:96-99 / 212-216, swin_attention_and_mlp.py:201-227). One warp owns one query row
// (N <= 256 keys: DeiT 197/198, Swin 49), values live in registers, reductions are warp shuffles.
#include "host_util.h"
#include "ofq_b200.h"
#include "ptx.cuh"
#include <cstdint>

namespace {

constexpr int kMaxPer = 8;  // keys per lane -> N <= 256

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint16_t to_bf16(float v) { return (uint16_t)(pack_bf16x2(v, 0.f) & 0xffff); }
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
template <bool F16>
__device__ __forceinline__ uint16_t to16(float v) { return (uint16_t)(pack16x2<F16>(v, 0.f) & 0xffff); }

__global__ void __launch_bounds__(256)
softmax_quant_kernel(const float* __restrict__ S, int nz, int N, long long ld, int H,
                     const float* __restrict__ bias, const float* __restrict__ mask, int nW,
                     const float* __restrict__ s_eff, float qhi, float* __restrict__ P,
                     int8_t* __restrict__ codes, long long ldq, float* __restrict__ rowsum) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gr = (long long)blockIdx.x * 8 + warp;   // global row = z * N + n
    if (gr >= (long long)nz * N) return;
    const int z = (int)(gr / N), n = (int)(gr - (long long)z * N);
    const int h = z % H, b = z / H;
    const float* srow = S + ((long long)z * N + n) * ld;
    const float* brow = bias ? bias + ((long long)h * N + n) * N : nullptr;
    const float* mrow = mask ? mask + ((long long)(b % nW) * N + n) * N : nullptr;
    float v[kMaxPer];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        const int d = lane + 32 * i;
        float t = -INFINITY;
        if (d < N) {
            t = __ldg(srow + d);
            if (brow) t += __ldg(brow + d);
            if (mrow) t += __ldg(mrow + d);
        }
        v[i] = t;
        m = fmaxf(m, t);
    }
    m = wmax(m);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        const int d = lane + 32 * i;
        v[i] = d < N ? expf(v[i] - m) : 0.f;
        sum += v[i];
    }
    sum = wsum(sum);
    const float s = __ldg(s_eff + n);
    float csum = 0.f;
    float* prow = P ? P + ((long long)z * N + n) * ld : nullptr;
    int8_t* crow = codes + ((long long)z * N + n) * ldq;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        const int d = lane + 32 * i;
        if (d < N) {
            const float p = __fdiv_rn(v[i], sum);
            const float q = rintf(fminf(fmaxf(__fdiv_rn(p, s), 0.f), qhi));
            if (prow) prow[d] = p;
            crow[d] = (int8_t)(int)q;
            csum += q;
        } else if (d < ldq) {
            crow[d] = 0;
        }
    }
    csum = wsum(csum);
    if (lane == 0 && rowsum) rowsum[(long long)z * N + n] = s * csum;
}

// Backward: block = 32 query rows of one (b, h); 8 warps x 4 rows.
template <bool F16>
__global__ void __launch_bounds__(256)
softmax_quant_bwd_kernel(const float* __restrict__ dPq, const float* __restrict__ P, int N, long long ld, int H,
                         const float* __restrict__ s_eff, float qhi, float alpha, float g_s,
                         const float* __restrict__ ca, int ca_per_head, const float* __restrict__ rb, int planes,
                         const float* __restrict__ scale4, int a_rowscale,
                         uint16_t* __restrict__ out_a, uint16_t* __restrict__ out_bt, long long ldo,
                         float* __restrict__ colsum, float* __restrict__ d_s, float* __restrict__ dS32) {
    extern __shared__ float tile[];           // [32][N + 1] dS * rb[n] for the transposed output
    __shared__ float csum_s[kMaxPer * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.y, h = z % H, b = z / H;
    // output slab of (b, plane p, h): index ((b * planes + p) * H + h); plane stride = H slabs
    const long long zo = ((long long)b * planes * H + h) * N;
    const long long plane = (long long)H * N * ldo;
    const int n0 = blockIdx.x * 32;
    const int pitch = N + 1;
    for (int i = threadIdx.x; i < kMaxPer * 32; i += blockDim.x) csum_s[i] = 0.f;
    __syncthreads();
    const float* cav = ca ? (ca_per_head ? ca + (long long)h * N : ca) : nullptr;
    // power-of-two range scales of the fp16 operands (ofq_absmax_scale): [0] out_a, [2] out_bt
    const float sc_a = scale4 ? __ldg(scale4 + 0) : 1.f;
    const float sc_b = scale4 ? __ldg(scale4 + 2) : 1.f;
    float colacc[kMaxPer];
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) colacc[i] = 0.f;

    for (int rr = 0; rr < 4; ++rr) {
        const int ln = warp * 4 + rr;
        const int n = n0 + ln;
        if (n >= N) {                                   // warp-uniform; keep the tile defined
            for (int d = lane; d < N; d += 32) tile[ln * pitch + d] = 0.f;
            continue;
        }
        const long long ro = ((long long)z * N + n) * ld;
        const float s = __ldg(s_eff + n);
        float p[kMaxPer], dp[kMaxPer];
        float dot = 0.f, dsp = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) {
            const int d = lane + 32 * i;
            p[i] = 0.f; dp[i] = 0.f;
            if (d < N) {
                p[i] = __ldg(P + ro + d);
                const float gq = __ldg(dPq + ro + d);
                const float v = __fdiv_rn(p[i], s);
                const bool inside = v <= qhi;           // v >= 0 always holds for probabilities
                const float q = rintf(fminf(v, qhi));
                dsp += gq * (inside ? (q - v) : q);
                dp[i] = inside ? gq : 0.f;
                dot += p[i] * dp[i];
            }
        }
        dot = wsum(dot);
        dsp = wsum(dsp);
        if (lane == 0 && d_s) atomicAdd(d_s + n, g_s * dsp);
        const float rb_raw = rb ? __ldg(rb + n) : 1.f;
        const float rbv = rb_raw * sc_b;
        const float sca_row = a_rowscale ? sc_a * rb_raw : sc_a;
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) {
            const int d = lane + 32 * i;
            if (d < N) {
                const float ds_raw = p[i] * (dp[i] - dot);   // gradient w.r.t. the (scaled) logits / additive bias
                const float ds = alpha * ds_raw;             // gradient w.r.t. the un-scaled q.k product
                colacc[i] += ds;
                if (out_a) {
                    const float va = ds * (cav ? __ldg(cav + d) : 1.f) * sca_row;
                    const uint16_t hi16 = to16<F16>(va);
                    uint16_t* dst = out_a + (zo + n) * ldo + d;
                    *dst = hi16;
                    if (!F16 && planes == 2) dst[plane] = to_bf16(va - __uint_as_float((uint32_t)hi16 << 16));
                }
                if (dS32) dS32[ro + d] = ds_raw;
                tile[ln * pitch + d] = ds * rbv;
            } else if (out_a && d < ldo) {
                out_a[(zo + n) * ldo + d] = 0;
                if (!F16 && planes == 2) out_a[(zo + n) * ldo + d + plane] = 0;
            }
        }
    }
    if (colsum) {
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i)
            if (lane + 32 * i < N) atomicAdd(&csum_s[lane + 32 * i], colacc[i]);
    }
    __syncthreads();
    if (colsum)
        for (int d = threadIdx.x; d < N; d += blockDim.x) atomicAdd(colsum + (long long)z * N + d, csum_s[d]);
    if (out_bt) {
        // out_bt[z][d][n0 .. n0+31]: 4 threads per key row d, 8 consecutive n each (16 bytes)
        for (int item = threadIdx.x; item < N * 4; item += blockDim.x) {
            const int d = item >> 2, no = (item & 3) * 8;
            if (n0 + no >= ldo) continue;
            float t8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) t8[e] = tile[(no + e) * pitch + d];
            uint4 pk;
            pk.x = pack16x2<F16>(t8[0], t8[1]);
            pk.y = pack16x2<F16>(t8[2], t8[3]);
            pk.z = pack16x2<F16>(t8[4], t8[5]);
            pk.w = pack16x2<F16>(t8[6], t8[7]);
            uint16_t* dst = out_bt + (zo + d) * ldo + n0 + no;
            *reinterpret_cast<uint4*>(dst) = pk;
            if (!F16 && planes == 2) {
                uint4 lo;
                lo.x = pack_bf16x2(t8[0] - __uint_as_float(pk.x << 16), t8[1] - __uint_as_float(pk.x & 0xffff0000u));
                lo.y = pack_bf16x2(t8[2] - __uint_as_float(pk.y << 16), t8[3] - __uint_as_float(pk.y & 0xffff0000u));
                lo.z = pack_bf16x2(t8[4] - __uint_as_float(pk.z << 16), t8[5] - __uint_as_float(pk.z & 0xffff0000u));
                lo.w = pack_bf16x2(t8[6] - __uint_as_float(pk.w << 16), t8[7] - __uint_as_float(pk.w & 0xffff0000u));
                *reinterpret_cast<uint4*>(dst + plane) = lo;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------- vectorised paths
// DeiT-sized rows (N = 197/198, pitch 200) are 16-byte aligned: a lane owns the float4 columns {lane, lane + 32} of a
// row (8 keys), so a row is two 16-byte loads per lane, probabilities leave as float4, codes as packed 32-bit words and
// the 16-bit gradient operand as 8-byte stores. The two divisions per element of the scalar path (p = e / sum and
// v = p / s) are replaced by per-row reciprocals:
//   p  : Markstein's sequence on a Newton-refined reciprocal of the row sum (q = e r; q += fma(-sum, q, e) r), which is the
//        correctly rounded quotient for these operands (e in [0, 1], sum in [1, 256]) up to the ulp-level tie class the
//        DESIGN documents for expf / summation order;
//   code: rint(p * rcp(s)) unless the product is within 2e-4 of a rounding boundary, where the IEEE quotient decides
//        (same argument as lsq_code_fast in quant.cu): codes stay bit-exact for the stored probabilities.
__device__ __forceinline__ float rcp_fast(float s) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(s));
    return r;
}
// exact rint(clamp(p / s, 0, qhi)) with a cheap quotient; *vq receives the quotient used for the decision
__device__ __forceinline__ float prob_code(float p, float s, float inv_s, float qhi, float* vq) {
    float v = __fmul_rn(p, inv_s);
    float r = rintf(v);
    const float dv = fabsf(v - r);
    if (dv > 0.4998f || (dv < 2e-4f && r == qhi)) {     // rounding boundary, or the clamp bound itself (STE mask)
        v = __fdiv_rn(p, s);
        r = rintf(v);
    }
    *vq = v;
    return fminf(r, qhi);
}
__device__ __forceinline__ uint32_t pack4_u8(float a, float b, float c, float d) {
    return (uint32_t)(int)a | ((uint32_t)(int)b << 8) | ((uint32_t)(int)c << 16) | ((uint32_t)(int)d << 24);
}

constexpr int kVecRows = 2;   // rows per warp in flight (4 x 16 B loads per lane)

__global__ void __launch_bounds__(256)
softmax_quant_vec_kernel(const float* __restrict__ S, long long rows, int N, long long ld,
                         const float* __restrict__ s_eff, float qhi, float* __restrict__ P,
                         int8_t* __restrict__ codes, long long ldq, float* __restrict__ rowsum,
                         uint16_t* __restrict__ codes16, int f16) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long r0 = gw * kVecRows;
    if (r0 >= rows) return;
    const int nv = (int)(ld >> 2), nq = (int)(ldq >> 2);
    const int j0 = lane, j1 = lane + 32;
    const float NEG = -INFINITY;
    float4 a[kVecRows][2];
#pragma unroll
    for (int u = 0; u < kVecRows; ++u) {
        const long long r = r0 + u;
        a[u][0] = a[u][1] = make_float4(NEG, NEG, NEG, NEG);
        if (r < rows) {
            const float4* sp = reinterpret_cast<const float4*>(S + r * ld);
            if (j0 < nv) a[u][0] = __ldg(sp + j0);
            if (j1 < nv) a[u][1] = __ldg(sp + j1);
        }
    }
#pragma unroll
    for (int u = 0; u < kVecRows; ++u) {
        const long long r = r0 + u;
        if (r >= rows) break;                      // warp-uniform
        const int n = (int)(r % N);
        float x[8] = {a[u][0].x, a[u][0].y, a[u][0].z, a[u][0].w, a[u][1].x, a[u][1].y, a[u][1].z, a[u][1].w};
        float m = NEG;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = 4 * (e < 4 ? j0 : j1) + (e & 3);
            if (col >= N) x[e] = NEG;
            m = fmaxf(m, x[e]);
        }
        m = wmax(m);
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = 4 * (e < 4 ? j0 : j1) + (e & 3);
            x[e] = col < N ? ofq::softmax_exp(x[e] - m) : 0.f;
            sum += x[e];
        }
        sum = wsum(sum);
        float rinv = rcp_fast(sum);
        rinv = fmaf(rinv, fmaf(-sum, rinv, 1.0f), rinv);         // Newton step: rinv = RN(1 / sum) for almost every sum
        const float s = __ldg(s_eff + n);
        const float inv_s = rcp_fast(s);
        float q[8], csum = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float p = __fmul_rn(x[e], rinv);
            p = fmaf(fmaf(-sum, p, x[e]), rinv, p);              // Markstein correction of the quotient
            x[e] = p;
            float vq;
            q[e] = prob_code(p, s, inv_s, qhi, &vq);
            csum += q[e];
        }
        if (P) {
            float4* pp = reinterpret_cast<float4*>(P + r * ld);
            if (j0 < nv) pp[j0] = make_float4(x[0], x[1], x[2], x[3]);
            if (j1 < nv) pp[j1] = make_float4(x[4], x[5], x[6], x[7]);
        }
        uint32_t* cp = reinterpret_cast<uint32_t*>(codes + r * ldq);
        if (j0 < nq) cp[j0] = pack4_u8(q[0], q[1], q[2], q[3]);     // columns >= N hold p = 0 -> code 0
        if (j1 < nq) cp[j1] = pack4_u8(q[4], q[5], q[6], q[7]);
        if (codes16) {     // exact 16-bit copy (same pitch, zero padding): the operand of the dV GEMM of the backward
            uint2* hp = reinterpret_cast<uint2*>(codes16 + r * ldq);
            if (j0 < nq) hp[j0] = f16 ? make_uint2(pack_f16x2(q[0], q[1]), pack_f16x2(q[2], q[3]))
                                      : make_uint2(pack_bf16x2(q[0], q[1]), pack_bf16x2(q[2], q[3]));
            if (j1 < nq) hp[j1] = f16 ? make_uint2(pack_f16x2(q[4], q[5]), pack_f16x2(q[6], q[7]))
                                      : make_uint2(pack_bf16x2(q[4], q[5]), pack_bf16x2(q[6], q[7]));
        }
        csum = wsum(csum);
        if (lane == 0 && rowsum) rowsum[r] = s * csum;
    }
}

// Backward, single 16-bit output (out_a only, one plane): block = 4 warps, one (b, h) slab per block, a warp walks rows
// warp, warp + 4, ... two at a time; per-key column sums stay in 8 registers per lane and are folded across the 4 warps
// once, so colsum is written without atomics.
// 6 CTAs per SM (<= 85 registers): the 768 slabs of a DeiT-S batch then fit in ONE wave of 888 slots (5 per SM would be
// 740 slots: a second wave with 28 CTAs that costs as much as the first).
template <bool F16>
__global__ void __launch_bounds__(128, 6)
softmax_quant_bwd_vec_kernel(const float* __restrict__ dPq, const float* __restrict__ P, int N, long long ld, int H,
                             const float* __restrict__ s_eff, float qhi, float alpha, float g_s,
                             const float* __restrict__ ca, int ca_per_head, const float* __restrict__ rb,
                             const float* __restrict__ scale4, int a_rowscale, uint16_t* __restrict__ out_a,
                             long long ldo, float* __restrict__ colsum, float* __restrict__ d_s,
                             float* __restrict__ dS32, float* __restrict__ ds_part) {
    __shared__ float fold[4][kMaxPer * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.x, h = z % H;
    const int nv = (int)(ld >> 2), no = (int)(ldo >> 2);
    const int j0 = lane, j1 = lane + 32;
    const float sc_a = scale4 ? __ldg(scale4 + 0) : 1.f;
    const float* cav = ca ? (ca_per_head ? ca + (long long)h * N : ca) : nullptr;
    float cv[8], colacc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int col = 4 * (e < 4 ? j0 : j1) + (e & 3);
        cv[e] = (cav && col < N) ? __ldg(cav + col) : 1.f;
        colacc[e] = 0.f;
    }
    for (int nb = warp * kVecRows; nb < N; nb += 4 * kVecRows) {
        float4 pv[kVecRows][2], gv[kVecRows][2];
#pragma unroll
        for (int u = 0; u < kVecRows; ++u) {
            const int n = nb + u;
            pv[u][0] = pv[u][1] = gv[u][0] = gv[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
                const long long ro = ((long long)z * N + n) * ld;
                const float4* pp = reinterpret_cast<const float4*>(P + ro);
                const float4* gp = reinterpret_cast<const float4*>(dPq + ro);
                if (j0 < nv) { pv[u][0] = __ldg(pp + j0); gv[u][0] = __ldg(gp + j0); }
                if (j1 < nv) { pv[u][1] = __ldg(pp + j1); gv[u][1] = __ldg(gp + j1); }
            }
        }
#pragma unroll
        for (int u = 0; u < kVecRows; ++u) {
            const int n = nb + u;
            if (n >= N) break;                     // warp-uniform
            const long long ro = ((long long)z * N + n) * ld;
            const float s = __ldg(s_eff + n);
            const float inv_s = rcp_fast(s);
            float p[8] = {pv[u][0].x, pv[u][0].y, pv[u][0].z, pv[u][0].w, pv[u][1].x, pv[u][1].y, pv[u][1].z, pv[u][1].w};
            float g[8] = {gv[u][0].x, gv[u][0].y, gv[u][0].z, gv[u][0].w, gv[u][1].x, gv[u][1].y, gv[u][1].z, gv[u][1].w};
            float dot = 0.f, dsp = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int col = 4 * (e < 4 ? j0 : j1) + (e & 3);
                if (col >= N) { p[e] = 0.f; g[e] = 0.f; }       // the pitch padding of P / dPq is not defined
                float v;
                const float q = prob_code(p[e], s, inv_s, qhi, &v);
                const bool inside = v <= qhi;                    // v >= 0 always holds for probabilities
                dsp += g[e] * (inside ? (q - v) : q);
                g[e] = inside ? g[e] : 0.f;
                dot += p[e] * g[e];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                dot += __shfl_xor_sync(0xffffffffu, dot, o);
                dsp += __shfl_xor_sync(0xffffffffu, dsp, o);
            }
            if (lane == 0) {
                // 768 slabs x 198 rows all target the same 198 scale gradients: with a partial buffer the per-row term is a
                // plain store, reduced over the slabs by ds_reduce_kernel (deterministic, no same-address atomics)
                if (ds_part) ds_part[(long long)z * N + n] = dsp;
                else if (d_s) atomicAdd(d_s + n, g_s * dsp);
            }
            const float rb_raw = rb ? __ldg(rb + n) : 1.f;
            const float sca_row = a_rowscale ? sc_a * rb_raw : sc_a;
            float raw[8], va[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                raw[e] = p[e] * (g[e] - dot);                    // gradient w.r.t. the (scaled) logits / additive bias
                const float ds = alpha * raw[e];                 // gradient w.r.t. the un-scaled q.k product
                colacc[e] += ds;
                va[e] = ds * cv[e] * sca_row;                    // columns >= N: p = 0 -> exact zeros in the padding
            }
            if (out_a) {
                uint2* op = reinterpret_cast<uint2*>(out_a + ((long long)z * N + n) * ldo);
                if (j0 < no) op[j0] = make_uint2(pack16x2<F16>(va[0], va[1]), pack16x2<F16>(va[2], va[3]));
                if (j1 < no) op[j1] = make_uint2(pack16x2<F16>(va[4], va[5]), pack16x2<F16>(va[6], va[7]));
            }
            if (dS32) {
                float4* dp = reinterpret_cast<float4*>(dS32 + ro);
                if (j0 < nv) dp[j0] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                if (j1 < nv) dp[j1] = make_float4(raw[4], raw[5], raw[6], raw[7]);
            }
        }
    }
    if (colsum) {
#pragma unroll
        for (int e = 0; e < 8; ++e) fold[warp][(e < 4 ? j0 : j1) * 4 + (e & 3)] = colacc[e];
        __syncthreads();
        for (int d = threadIdx.x; d < N; d += blockDim.x)
            colsum[(long long)z * N + d] = (fold[0][d] + fold[1][d]) + (fold[2][d] + fold[3][d]);
    }
}

// d_s[n] = g_s * sum_z part[z][n]: block = 32 rows n x 8 slices of z (fixed order)
__global__ void __launch_bounds__(256)
ds_reduce_kernel(const float* __restrict__ part, int nz, int N, float g_s, float* __restrict__ d_s) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (n < N)
        for (int z = ty; z < nz; z += 8) acc += part[(long long)z * N + n];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][tx];
        d_s[n] = g_s * s;
    }
}

}  // namespace

extern "C" int ofq_softmax_quant_ex(const float* S, int nz, int N, long long ld, int H, const float* bias,
                                    const float* mask, int nW, const float* s_eff, int qhi, float* P, int8_t* codes,
                                    long long ldq, float* rowsum, void* codes16, int fmt16, void* stream) {
    OFQ_REQUIRE(S && s_eff && codes && nz > 0 && N > 0 && H > 0, "ofq_softmax_quant: bad argument");
    OFQ_REQUIRE(N <= kMaxPer * 32, "ofq_softmax_quant: at most 256 keys per row are supported");
    OFQ_REQUIRE(ld >= N && ldq >= N && qhi > 0 && qhi <= 127, "ofq_softmax_quant: bad pitch or level count");
    OFQ_REQUIRE(!mask || nW > 0, "ofq_softmax_quant: mask needs nW");
    OFQ_CHECK_ARCH();
    const long long rows = (long long)nz * N;
    const bool vec = !bias && !mask && ld % 4 == 0 && ldq % 4 == 0 && (uintptr_t)S % 16 == 0 && (!P || (uintptr_t)P % 16 == 0) &&
                     (uintptr_t)codes % 4 == 0 && (!codes16 || (uintptr_t)codes16 % 8 == 0);
    OFQ_REQUIRE(!codes16 || vec, "ofq_softmax_quant: the 16-bit copy is produced by the vectorised path only "
                                 "(no bias / mask, pitches that are multiples of 4, aligned pointers)");
    OFQ_REQUIRE(!codes16 || fmt16 == OFQ_FMT_BF16 || fmt16 == OFQ_FMT_F16, "ofq_softmax_quant: bad 16-bit format");
    if (vec) {
        softmax_quant_vec_kernel<<<(unsigned)((rows + 8 * kVecRows - 1) / (8 * kVecRows)), 256, 0, (cudaStream_t)stream>>>(
            S, rows, N, ld, s_eff, (float)qhi, P, codes, ldq, rowsum, (uint16_t*)codes16, fmt16 == OFQ_FMT_F16);
        OFQ_CUDA(cudaGetLastError());
        return 0;
    }
    softmax_quant_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        S, nz, N, ld, H, bias, mask, nW > 0 ? nW : 1, s_eff, (float)qhi, P, codes, ldq, rowsum);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_softmax_quant(const float* S, int nz, int N, long long ld, int H, const float* bias,
                                 const float* mask, int nW, const float* s_eff, int qhi, float* P, int8_t* codes,
                                 long long ldq, float* rowsum, void* stream) {
    return ofq_softmax_quant_ex(S, nz, N, ld, H, bias, mask, nW, s_eff, qhi, P, codes, ldq, rowsum, nullptr, OFQ_FMT_F16, stream);
}

extern "C" int ofq_softmax_quant_bwd_ex(const float* dPq, const float* P, int nz, int N, long long ld, int H,
                                     const float* s_eff, int qhi, float alpha, float g_s, const float* ca,
                                     int ca_per_head, const float* rb, int planes, void* out_a, void* out_bt,
                                     long long ldo, float* colsum, float* d_s, float* dS32, int out_fmt,
                                     const float* scale4, int a_rowscale, float* ds_partial, void* stream) {
    OFQ_REQUIRE(dPq && P && s_eff && nz > 0 && N > 0 && H > 0, "ofq_softmax_quant_bwd: bad argument");
    OFQ_REQUIRE(N <= kMaxPer * 32, "ofq_softmax_quant_bwd: at most 256 keys per row are supported");
    OFQ_REQUIRE((!out_a && !out_bt) || (ldo % 8 == 0 && ldo >= N), "ofq_softmax_quant_bwd: output pitch must be a multiple of 8 and >= N");
    OFQ_REQUIRE(!out_bt || (uintptr_t)out_bt % 16 == 0, "ofq_softmax_quant_bwd: out_bt alignment");
    OFQ_REQUIRE(planes == 1 || planes == 2, "ofq_softmax_quant_bwd: planes must be 1 or 2");
    OFQ_REQUIRE(out_fmt == OFQ_FMT_BF16 || (out_fmt == OFQ_FMT_F16 && planes == 1), "ofq_softmax_quant_bwd: fp16 output is single-plane");
    OFQ_CHECK_ARCH();
    // single-output, one-plane requests on 16-byte aligned rows take the vectorised kernel (colsum is then written,
    // not accumulated: pre-zeroing it is harmless)
    const bool vec = !out_bt && planes == 1 && ld % 4 == 0 && ldo % 4 == 0 && (uintptr_t)dPq % 16 == 0 && (uintptr_t)P % 16 == 0 &&
                     (!out_a || (uintptr_t)out_a % 8 == 0) && (!dS32 || (uintptr_t)dS32 % 16 == 0);
    if (vec) {
        if (out_fmt == OFQ_FMT_F16)
            softmax_quant_bwd_vec_kernel<true><<<nz, 128, 0, (cudaStream_t)stream>>>(
                dPq, P, N, ld, H, s_eff, (float)qhi, alpha, g_s, ca, ca_per_head, rb, scale4, a_rowscale, (uint16_t*)out_a, ldo,
                colsum, d_s, dS32, d_s ? ds_partial : nullptr);
        else
            softmax_quant_bwd_vec_kernel<false><<<nz, 128, 0, (cudaStream_t)stream>>>(
                dPq, P, N, ld, H, s_eff, (float)qhi, alpha, g_s, ca, ca_per_head, rb, scale4, a_rowscale, (uint16_t*)out_a, ldo,
                colsum, d_s, dS32, d_s ? ds_partial : nullptr);
        if (d_s && ds_partial)
            ds_reduce_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(ds_partial, nz, N, g_s, d_s);
        OFQ_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 grid((N + 31) / 32, nz);
    const size_t smem = (size_t)32 * (N + 1) * sizeof(float);
    if (out_fmt == OFQ_FMT_F16)
        softmax_quant_bwd_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(
            dPq, P, N, ld, H, s_eff, (float)qhi, alpha, g_s, ca, ca_per_head, rb, planes, scale4, a_rowscale, (uint16_t*)out_a,
            (uint16_t*)out_bt, ldo, colsum, d_s, dS32);
    else
        softmax_quant_bwd_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(
            dPq, P, N, ld, H, s_eff, (float)qhi, alpha, g_s, ca, ca_per_head, rb, planes, scale4, a_rowscale, (uint16_t*)out_a,
            (uint16_t*)out_bt, ldo, colsum, d_s, dS32);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ofq_softmax_quant_bwd(const float* dPq, const float* P, int nz, int N, long long ld, int H,
                                     const float* s_eff, int qhi, float alpha, float g_s, const float* ca,
                                     int ca_per_head, const float* rb, int planes, void* out_a, void* out_bt,
                                     long long ldo, float* colsum, float* d_s, float* dS32, int out_fmt,
                                     const float* scale4, int a_rowscale, void* stream) {
    return ofq_softmax_quant_bwd_ex(dPq, P, nz, N, ld, H, s_eff, qhi, alpha, g_s, ca, ca_per_head, rb, planes, out_a, out_bt, ldo,
                                    colsum, d_s, dS32, out_fmt, scale4, a_rowscale, nullptr, stream);
}
