/* ofq_b200 — C-ABI of the B200-native OFQ quantization-aware-training hot path.
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t (as void*); nothing here knows
 * about torch. All functions return 0 on success or a negative OFQ_ERR_* code; ofq_last_error() gives the
 * message (thread-local). There is no CPU fallback: on a device that is not sm_100 every compute entry
 * point fails with OFQ_ERR_ARCH.
 *
 * "Reference" citations are file:line in nbasyl/OFQ (the tree mounted at /root/reference during development).
 * The Python host side (ofq_b200/quantization/...) mirrors the reference's nn.Module API on top of this ABI;
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 */
#ifndef OFQ_B200_H
#define OFQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFQ_API __attribute__((visibility("default")))

#define OFQ_ERR_ARG   (-1)  /* bad argument (shape / alignment / null) */
#define OFQ_ERR_CUDA  (-2)  /* CUDA runtime / driver error */
#define OFQ_ERR_ARCH  (-3)  /* current device is not sm_100 (B200) */

OFQ_API int ofq_version(void);
OFQ_API const char* ofq_last_error(void);
/* 1 if the current CUDA device is compute capability 10.x, 0 otherwise, <0 on CUDA error. */
OFQ_API int ofq_device_ok(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM engine (tcgen05 + TMEM + TMA).  Replaces F.linear / einsum / bmm / @ on fake-quant
 * floats in the reference (qlinear.py:69, attention.py:96,102,180,200,210,219,
 * swin_attention_and_mlp.py:201,229) and the autograd backward of those calls.
 *
 *   D[b][m][n] (=|+=) (sum_{k2,k} A[b][k2][m][k] * B[b][k2][n][k]) * rs[m] * cs[n] + rt[m] * ct[n]
 *
 * Both operands are K-major (k contiguous). kind I8: int8 codes, exact int32 accumulation.
 * kind BF16: bf16 operands, fp32 accumulation. Strides are in elements; a batch stride of 0 means the
 * operand is shared across that batch axis. Row/k2/batch strides must be multiples of 16 bytes.
 * NULL vectors read as 1; the rank-1 term is added only when rt or ct is non-NULL.
 * With splits > 1 (split-K) the output must be pre-zeroed and `accumulate` set (fp32 atomics).
 */
#define OFQ_GEMM_I8   0
#define OFQ_GEMM_BF16 1

typedef struct {
    const void* ptr;
    long long row_stride; /* elements between consecutive rows (m or n) */
    long long k2_stride;  /* elements between outer-K slices (0 if k2 == 1) */
    long long bstride1;   /* elements between batch-axis-1 entries, 0 = broadcast */
    long long bstride2;   /* elements between batch-axis-2 entries, 0 = broadcast */
} ofq_operand_t;

typedef struct {
    const float* ptr;     /* NULL = all ones */
    int period;           /* value for index i is ptr[i % period]; <= 0 means no wrap */
    long long bstride1;   /* offset per batch-axis-1 index */
    long long bstride2;   /* offset per batch-axis-2 index */
} ofq_vec_t;

typedef struct {
    float* ptr;
    long long ld;         /* elements between consecutive output rows */
    long long bstride1;
    long long bstride2;
    int accumulate;       /* 0: store, 1: atomic add into existing contents */
} ofq_gemm_out_t;

OFQ_API int ofq_gemm(int kind, const ofq_operand_t* A, const ofq_operand_t* B, const ofq_gemm_out_t* out,
             int M, int N, int K, int k2, int nb1, int nb2, int splits,
             const ofq_vec_t* rs, const ofq_vec_t* cs, const ofq_vec_t* rt, const ofq_vec_t* ct,
             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFQ_B200_H */
