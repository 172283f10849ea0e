// K4: query-key reparameterisation weight product W_qk[h] = W_q[h]^T W_k[h] (attention.py:190-194) and its
// backward, in true fp32 (the product feeds StatsQ, so no reduced-precision tensor-core path here).
// A small register-tiled batched SGEMM with arbitrary element strides: weights only, 0.68 GMAC per step.
#include "host_util.h"
#include "ofq_b200.h"

namespace {

// C[b][m][n] = sum_k A[b][k*sAk + m*sAm] * B[b][k*sBk + n*sBn];  64x64 tile, 256 threads, 4x4 per thread.
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long long sAk, long long sAm, long long bsA,
                     const float* __restrict__ B, long long sBk, long long sBn, long long bsB,
                     float* __restrict__ C, long long ldc, long long bsC, int M, int N, int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const float* Ab = A + b * bsA;
    const float* Bb = B + b * bsB;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = t; i < 16 * 64; i += 256) {
            int kk, mm;
            if (sAm == 1) { kk = i >> 6; mm = i & 63; } else { kk = i & 15; mm = i >> 4; }
            const int k = k0 + kk, m = m0 + mm;
            As[kk][mm] = (k < K && m < M) ? __ldg(Ab + k * sAk + m * sAm) : 0.f;
            int kb, nn;
            if (sBn == 1) { kb = i >> 6; nn = i & 63; } else { kb = i & 15; nn = i >> 4; }
            const int k2 = k0 + kb, n = n0 + nn;
            Bs[kb][nn] = (k2 < K && n < N) ? __ldg(Bb + k2 * sBk + n * sBn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; bb[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Cb = C + b * bsC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) Cb[m * ldc + n] = acc[i][j];
        }
    }
}

int launch(const float* A, long long sAk, long long sAm, long long bsA, const float* B, long long sBk, long long sBn,
           long long bsB, float* C, long long ldc, long long bsC, int M, int N, int K, int nb, cudaStream_t st) {
    dim3 grid((N + 63) / 64, (M + 63) / 64, nb);
    sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, sAk, sAm, bsA, B, sBk, sBn, bsB, C, ldc, bsC, M, N, K);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" int ofq_wqk_compose(const float* wq, const float* wk, int H, int hd, int C, float* wqk, void* stream) {
    OFQ_REQUIRE(wq && wk && wqk && H > 0 && hd > 0 && C > 0, "ofq_wqk_compose: bad argument");
    OFQ_CHECK_ARCH();
    // wqk[h][i][j] = sum_d wq[h*hd+d][i] * wk[h*hd+d][j]
    return launch(wq, C, 1, (long long)hd * C, wk, C, 1, (long long)hd * C, wqk, C, (long long)C * C, C, C, hd, H,
                  (cudaStream_t)stream);
}

extern "C" int ofq_wqk_compose_bwd(const float* dwqk, const float* wq, const float* wk, int H, int hd, int C,
                                   float* dwq, float* dwk, void* stream) {
    OFQ_REQUIRE(dwqk && wq && wk && dwq && dwk && H > 0 && hd > 0 && C > 0, "ofq_wqk_compose_bwd: bad argument");
    OFQ_CHECK_ARCH();
    cudaStream_t st = (cudaStream_t)stream;
    // dwq[h*hd+d][i] = sum_j wk[h*hd+d][j] * dwqk[h][i][j]
    int rc = launch(wk, 1, C, (long long)hd * C, dwqk, 1, C, (long long)C * C, dwq, C, (long long)hd * C, hd, C, C, H, st);
    if (rc) return rc;
    // dwk[h*hd+d][j] = sum_i wq[h*hd+d][i] * dwqk[h][i][j]
    return launch(wq, 1, C, (long long)hd * C, dwqk, C, 1, (long long)C * C, dwk, C, (long long)hd * C, hd, C, C, H, st);
}
