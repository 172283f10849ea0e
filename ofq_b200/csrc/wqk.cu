// K4: query-key reparameterisation weight product W_qk[h] = W_q[h]^T W_k[h] (attention.py:190-194) and its
// backward, in true fp32 (the product feeds StatsQ, so no reduced-precision tensor-core path here).
// A small register-tiled batched SGEMM with arbitrary element strides: weights only, 0.68 GMAC per step.
#include "host_util.h"
#include "ofq_b200.h"

namespace {

// C[b][m][n] = sum_k A[b][k*sAk + m*sAm] * B[b][k*sBk + n*sBn];  TM x 64 tile, 256 threads, (TM/16) x 4 per thread.
// Up to two independent problems share one launch (blockIdx.z = problem * nb + batch): the two backward products
// of a layer are 36 + 36 tiles of 64 x 64, far fewer than the 148 SMs, so they run side by side on 16-row tiles (288 CTAs).
struct SgemmProblem {
    const float* A; long long sAk, sAm, bsA;
    const float* B; long long sBk, sBn, bsB;
    float* C; long long ldc, bsC;
};
struct SgemmPair { SgemmProblem p[2]; };

// BK: k-extent of a shared-memory tile. The loop is not double buffered, so every tile exposes one global-load round trip:
// the small backward products (16-row tiles, K = 384) take 64-deep tiles (6 round trips instead of 24, 20 loads in flight
// per thread).
template <int TM, int BK>
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const SgemmPair pair, int nb, int M, int N, int K) {
    constexpr int RM = TM / 16;
    __shared__ float As[BK][TM + 4];
    __shared__ float Bs[BK][64 + 4];
    const SgemmProblem& q = pair.p[blockIdx.z / nb];
    const int b = blockIdx.z % nb;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * 64;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const float* Ab = q.A + b * q.bsA;
    const float* Bb = q.B + b * q.bsB;
    const long long sAk = q.sAk, sAm = q.sAm, sBk = q.sBk, sBn = q.sBn;
    float acc[RM][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll 4
        for (int i = t; i < BK * TM; i += 256) {
            int kk, mm;
            if (sAm == 1) { kk = i / TM; mm = i % TM; } else { kk = i % BK; mm = i / BK; }
            const int k = k0 + kk, m = m0 + mm;
            As[kk][mm] = (k < K && m < M) ? __ldg(Ab + k * sAk + m * sAm) : 0.f;
        }
#pragma unroll 4
        for (int i = t; i < BK * 64; i += 256) {
            int kb, nn;
            if (sBn == 1) { kb = i >> 6; nn = i & 63; } else { kb = i % BK; nn = i / BK; }
            const int k2 = k0 + kb, n = n0 + nn;
            Bs[kb][nn] = (k2 < K && n < N) ? __ldg(Bb + k2 * sBk + n * sBn) : 0.f;
        }
        __syncthreads();
#pragma unroll 16
        for (int kk = 0; kk < BK; ++kk) {
            float a[RM], bb[4];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w;
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Cb = q.C + b * q.bsC;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty * RM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) Cb[m * q.ldc + n] = acc[i][j];
        }
    }
}

// W_qk of every layer in one launch: blockIdx.z = job * H + head, jobs = {wq, wk, wqk} pointer triples in device memory
struct WqkJob { const float* wq; const float* wk; float* wqk; };
template <int TM>
__global__ void __launch_bounds__(256)
wqk_compose_multi_kernel(const WqkJob* __restrict__ table, int nb, int hd, int C) {
    constexpr int RM = TM / 16;
    __shared__ float As[16][TM + 4];
    __shared__ float Bs[16][64 + 4];
    const WqkJob job = table[blockIdx.z / nb];
    // wqk[h][i][j] = sum_d wq[h*hd+d][i] * wk[h*hd+d][j]
    const SgemmProblem q = {job.wq, C, 1, (long long)hd * C, job.wk, C, 1, (long long)hd * C, job.wqk, C, (long long)C * C};
    const int M = C, N = C, K = hd;
    const int b = blockIdx.z % nb;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * 64;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const float* Ab = q.A + b * q.bsA;
    const float* Bb = q.B + b * q.bsB;
    const long long sAk = q.sAk, sAm = q.sAm, sBk = q.sBk, sBn = q.sBn;
    float acc[RM][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = t; i < 16 * TM; i += 256) {
            int kk, mm;
            if (sAm == 1) { kk = i / TM; mm = i % TM; } else { kk = i & 15; mm = i >> 4; }
            const int k = k0 + kk, m = m0 + mm;
            As[kk][mm] = (k < K && m < M) ? __ldg(Ab + k * sAk + m * sAm) : 0.f;
        }
        for (int i = t; i < 16 * 64; i += 256) {
            int kb, nn;
            if (sBn == 1) { kb = i >> 6; nn = i & 63; } else { kb = i & 15; nn = i >> 4; }
            const int k2 = k0 + kb, n = n0 + nn;
            Bs[kb][nn] = (k2 < K && n < N) ? __ldg(Bb + k2 * sBk + n * sBn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[RM], bb[4];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Cb = q.C + b * q.bsC;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty * RM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) Cb[m * q.ldc + n] = acc[i][j];
        }
    }
}

int launch(const SgemmPair& pair, int nprob, int M, int N, int K, int nb, cudaStream_t st) {
    const long long tiles64 = (long long)((N + 63) / 64) * ((M + 63) / 64) * nb * nprob;
    const long long tiles32 = (long long)((N + 63) / 64) * ((M + 31) / 32) * nb * nprob;
    if (tiles64 >= 148) {
        dim3 grid((N + 63) / 64, (M + 63) / 64, nb * nprob);
        sgemm_strided_kernel<64, 16><<<grid, 256, 0, st>>>(pair, nb, M, N, K);
    } else if (false && tiles32 >= 128) {   // measured slower than the 16-row tiles on B200 (67 vs 58 us): kept for reference   // the backward pair of a DeiT-S layer: 144 CTAs of 32 x 64 (2 x 4 outputs per thread)
        dim3 grid((N + 63) / 64, (M + 31) / 32, nb * nprob);
        sgemm_strided_kernel<32, 32><<<grid, 256, 0, st>>>(pair, nb, M, N, K);
    } else {
        dim3 grid((N + 63) / 64, (M + 15) / 16, nb * nprob);
        sgemm_strided_kernel<16, 64><<<grid, 256, 0, st>>>(pair, nb, M, N, K);
    }
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" int ofq_wqk_compose(const float* wq, const float* wk, int H, int hd, int C, float* wqk, void* stream) {
    OFQ_REQUIRE(wq && wk && wqk && H > 0 && hd > 0 && C > 0, "ofq_wqk_compose: bad argument");
    OFQ_CHECK_ARCH();
    // wqk[h][i][j] = sum_d wq[h*hd+d][i] * wk[h*hd+d][j]
    SgemmPair pair = {};
    pair.p[0] = {wq, C, 1, (long long)hd * C, wk, C, 1, (long long)hd * C, wqk, C, (long long)C * C};
    return launch(pair, 1, C, C, hd, H, (cudaStream_t)stream);
}

extern "C" int ofq_wqk_compose_bwd(const float* dwqk, const float* wq, const float* wk, int H, int hd, int C,
                                   float* dwq, float* dwk, void* stream) {
    OFQ_REQUIRE(dwqk && wq && wk && dwq && dwk && H > 0 && hd > 0 && C > 0, "ofq_wqk_compose_bwd: bad argument");
    OFQ_CHECK_ARCH();
    SgemmPair pair;
    // dwq[h*hd+d][i] = sum_j wk[h*hd+d][j] * dwqk[h][i][j]
    pair.p[0] = {wk, 1, C, (long long)hd * C, dwqk, 1, C, (long long)C * C, dwq, C, (long long)hd * C};
    // dwk[h*hd+d][j] = sum_i wq[h*hd+d][i] * dwqk[h][i][j]
    pair.p[1] = {wq, 1, C, (long long)hd * C, dwqk, C, 1, (long long)C * C, dwk, C, (long long)hd * C};
    return launch(pair, 2, hd, C, C, H, (cudaStream_t)stream);
}

// Every layer's W_qk in one launch. table: device array of n_jobs records {wq, wk, wqk} (three pointers, 24 bytes).
extern "C" int ofq_wqk_compose_multi(const void* table, int n_jobs, int H, int hd, int C, void* stream) {
    OFQ_REQUIRE(table && n_jobs > 0 && H > 0 && hd > 0 && C > 0, "ofq_wqk_compose_multi: bad argument");
    OFQ_REQUIRE((long long)n_jobs * H <= 65535, "ofq_wqk_compose_multi: too many (layer, head) pairs");
    static_assert(sizeof(WqkJob) == 24, "WqkJob layout is part of the C-ABI");
    OFQ_CHECK_ARCH();
    dim3 grid((C + 63) / 64, (C + 63) / 64, n_jobs * H);
    wqk_compose_multi_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>((const WqkJob*)table, H, hd, C);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}
