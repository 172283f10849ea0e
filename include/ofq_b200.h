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
 * kind BF16 / F16: bf16 / fp16 operands, fp32 accumulation. Strides are in elements; a batch stride of 0 means the
 * operand is shared across that batch axis. Row/k2/batch strides must be multiples of 16 bytes.
 * NULL vectors read as 1; the rank-1 term is added only when rt or ct is non-NULL.
 * With splits > 1 (split-K) the output must be pre-zeroed and `accumulate` set (fp32 atomics). splits = 0: the library
 * chooses (split-K only for accumulating outputs), from the tile count, the K extent and the number of SMs.
 */
#define OFQ_GEMM_I8   0
#define OFQ_GEMM_BF16 1
#define OFQ_GEMM_F16  2   /* fp16 operands, fp32 accumulation (range-scaled single-plane backward operands) */

/* 16-bit operand formats written by the gradient-preparation kernels */
#define OFQ_FMT_BF16 0
#define OFQ_FMT_F16  1

typedef struct {
    const void* ptr;
    long long row_stride; /* elements between consecutive rows (m or n) */
    long long k2_stride;  /* elements between outer-K slices; 0 = the operand is shared by all slices */
    int k2_mod;           /* > 0: slice index used for this operand is (k2 index % k2_mod); 0: the k2 index itself */
    int dual_delta;       /* A operand only, bf16: > 0 loads slices k2 and k2 + dual_delta in the same pipeline stage and
                             multiplies both with one copy of B (hi / lo planes of a gradient operand); the operand then
                             holds k2 + dual_delta slices. 0 = off */
    long long bstride1;   /* elements between batch-axis-1 entries, 0 = broadcast */
    long long bstride2;   /* elements between batch-axis-2 entries, 0 = broadcast */
    int mn_major;         /* 16-bit kinds only: 1 = the operand is stored [k][row] (rows contiguous) and row_stride is the
                             distance between consecutive k; lets a row-major [tokens][features] tensor serve as the
                             transposed operand of dW = dY^T X without a transposed copy. 0 = K-major */
} ofq_operand_t;

typedef struct {
    const float* ptr;     /* NULL = all ones */
    int period;           /* value for index i is ptr[i % period]; <= 0 means no wrap (honoured by rs, rt and cs) */
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

/* Same; out_absmax (optional, one float PRE-SET to 0) receives max |D| over the whole problem, tracked by the epilogue
 * (atomicMax on the float's bits): the bound from which the fp16 range scale of the next gradient operand is derived
 * without a pass over D (ofq_scale_from_max). Plain stores only (no accumulate, no split-K). */
OFQ_API int ofq_gemm_ex(int kind, const ofq_operand_t* A, const ofq_operand_t* B, const ofq_gemm_out_t* out,
             int M, int N, int K, int k2, int nb1, int nb2, int splits,
             const ofq_vec_t* rs, const ofq_vec_t* cs, const ofq_vec_t* rt, const ofq_vec_t* ct,
             float* out_absmax, void* stream);

/* int8 GEMM whose epilogue IS the LSQ quantizer of its output (reference attention.py:200-207: qkx = Linear(x_hat) ->
 * move_qkx_b4 -> LsqQuantizer): y = acc * rs[m] * cs[n] + rt[m] * ct[n] exactly as ofq_gemm forms it, then
 *   v = (y + b4[n]) / s,  s = s_eff[(m % period) * nseg + n / seg_len],  q = rint(clamp(v, qlo, qhi))
 * with the same guarded quotient as ofq_lsq_quant (codes bit-identical to ofq_gemm followed by ofq_lsq_quant_ex).
 * The fp32 product is never written. Outputs (each optional except codes):
 *   codes   int8 [M, N]                 the operand of the next integer GEMM
 *   codes16 fp16 / bf16 [M, N]          exact copy for the 16-bit backward GEMMs
 *   res16   fp16 [M, N]                 what the quantizer's backward needs instead of y: q - v where qlo <= v <= qhi (the
 *                                       straight-through region; v = (y + b4) * inv_s as ofq_lsq_bwd evaluates it), -2 where v < qlo,
 *                                       +2 where v > qhi (|q - v| <= 1/2, so the three cases cannot be confused)
 *   rowdot  fp32 [M, nseg]              sum over the segment's columns of dot_u[n] * q (the column term of the attention logits)
 * seg_len % 32 == 0; N % 16 == 0; leading dimensions multiples of 16 elements. workspace: M * ceil(N / 32) floats when
 * rowdot is requested. */
typedef struct {
    int8_t* codes;      long long ld_codes;
    void* codes16;      long long ld_codes16;  int fmt16;     /* OFQ_FMT_F16 / OFQ_FMT_BF16 */
    void* res16;        long long ld_res16;
    const float* b4;        /* [N] or NULL */
    const float* s_eff;     /* [period * nseg] */
    const float* inv_s;     /* [period * nseg], 1.0f / s_eff */
    int period, nseg, seg_len;
    float qlo, qhi;
    const float* dot_u;     /* [N] or NULL */
    float* rowdot;          /* [M, nseg] or NULL */
    float* workspace;
} ofq_gemm_lsq_t;

OFQ_API int ofq_gemm_lsq(const ofq_operand_t* A, const ofq_operand_t* B, int M, int N, int K,
                         const ofq_vec_t* rs, const ofq_vec_t* cs, const ofq_vec_t* rt, const ofq_vec_t* ct,
                         const ofq_gemm_lsq_t* q, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1  StatsQuantizer.forward as integer codes (reference statsq.py:133-150).
 *   sf[r]       = 2 * mean_c |w[r][c]|                       (row sum accumulated in fp64, then fp32 ops)
 *   k           = rint( clamp(w/sf, -1, 1-1e-6) * n - 0.5 ),  n = 2^(bits-1)       (IEEE div, no FMA)
 *   codes[r][c] = 2k+1  (odd, |code| <= 2^bits - 1)  so that  w_q = sf/(2n) * code
 *   colscale[r] = sf[r] / (2n)
 *   colterm[r]  = colscale[r] * sum_c aft[c] * codes[r][c] + bias[r]   (optional: the move_aft shift of the
 *                 layer input folded through the weights, qlinear.py:68-71; aft/bias may be NULL)
 *   inv_colscale= optional 1 / colscale[r] (epilogue un-scale vector of the backward dW GEMM)
 *   kminmax     = optional int32[2] {min k, max k} over the whole tensor, atomically updated (pre-set to
 *                 {INT_MAX, INT_MIN}); cga.py:459-463 needs it.
 */
OFQ_API int ofq_statsq_codes(const float* w, int rows, int cols, long long ldw, int bits, int8_t* codes,
                             long long ldq, float* colscale, float* sf, const float* aft, const float* bias,
                             float* colterm, int* kminmax, float* inv_colscale, void* stream);
/* Same, plus an optional exact 16-bit copy codes16 [rows][cols] (OFQ_FMT_BF16 / _F16) of the codes: the B operand of
 * the layer's dX GEMM in the backward, written while the codes are in registers. */
OFQ_API int ofq_statsq_codes_ex(const float* w, int rows, int cols, long long ldw, int bits, int8_t* codes,
                                long long ldq, float* colscale, float* sf, const float* aft, const float* bias,
                                float* colterm, int* kminmax, float* inv_colscale, void* codes16, int fmt16, void* stream);

/* LSQ effective step size (lsq.py:593): out[i] = (a - a*g) + a*g with a = alpha[i] > 1e-5 ? alpha[i] : 1e-5,
 * evaluated in fp32 exactly as grad_scale(clip(alpha)) does. out_recip (optional) receives 1 / out[i]. */
OFQ_API int ofq_lsq_effective_scale(const float* alpha, int n, float g, float* out, float* out_recip, void* stream);

/* K2  LearnableBias + LsqQuantizer / LsqQuantizer4v forward as integer codes
 * (qbias.py:10-13, lsq.py:571-602, 757-790):   codes = rint( clamp( (x + b4[col]) / s_eff, qlo, qhi ) ).
 *   x        [rows][cols] fp32, row stride ldx; a row is `nseg` segments of cols/nseg columns
 *            (nseg = heads for the (B, N*H, C) view of qkx, attention.py:202-205; 1 otherwise)
 *   b4       shift of length cols
 *   s_eff    effective scales from ofq_lsq_effective_scale;
 *            scale_mode OFQ_SCALE_PER_ROW: index (row % period) * nseg + segment   (per token [, head])
 *            scale_mode OFQ_SCALE_PER_COL: index col                               (LsqQuantizer4v)
 */
#define OFQ_SCALE_PER_ROW 0
#define OFQ_SCALE_PER_COL 1
#define OFQ_ACT_NONE 0
#define OFQ_ACT_GELU 1   /* nn.GELU() (erf form) applied to x before the shift: the fc2 input of QMLP (qlinear.py:123-136) */
#define OFQ_ACT_RES16 2   /* ofq_lsq_bwd_ex only: `x` is the fp16 residual plane written by ofq_gemm_lsq (pitch ldx in halves), not the input */
OFQ_API int ofq_lsq_quant(const float* x, long long rows, int cols, long long ldx, const float* b4,
                          const float* s_eff, int scale_mode, int period, int nseg, int qlo, int qhi,
                          int8_t* codes, long long ldq, void* stream);

/* Same with (a) an activation fused in front of the quantizer, codes = Q(act(x) + b4), so that the GELU output of QMLP
 * never travels through HBM, and (b) an optional exact 16-bit copy codes16 [rows][ld16] (fmt16 = OFQ_FMT_BF16 / _F16)
 * of the codes, written in the same pass: the operand the backward GEMMs read, and (c) optional segment-wise dot
 * products of the codes with a vector dot_u[cols] (ofq_codes_rowdot in the same pass): the share of 128-column group j of
 * a segment goes to dot_part[j][row * nseg + seg] (cols/nseg/128 planes, to be summed by the caller; needs 128-column
 * groups that do not straddle segments). */
OFQ_API int ofq_lsq_quant_ex(const float* x, long long rows, int cols, long long ldx, const float* b4,
                             const float* s_eff, int scale_mode, int period, int nseg, int qlo, int qhi, int act,
                             int8_t* codes, long long ldq, void* codes16, long long ld16, int fmt16,
                             const float* dot_u, float* dot_part, void* stream);

/* Backward of (LearnableBias -> LSQ -> LearnableBias) given dy = dL/d(x_hat) (autograd of lsq.py:571-602):
 *   v = (x + b4)/s_eff;  inside = qlo <= v <= qhi;  q = rint(clamp(v))
 *   dx = dy * inside;  d_aft[c] = sum dy;  d_b4[c] = sum dx;  d_s[idx] = g * sum dy * (q - inside * v)
 * Partial sums go to `workspace` (float[ofq_lsq_bwd_workspace(...)], no atomics, deterministic) and are
 * reduced by ofq_lsq_bwd_finalize. dx may alias dy.
 * zero_sum != 0 declares that sum_rows dy is analytically zero for every column (the K operand of attention: a shift of
 * K only adds a term that is constant along the softmax axis); d_b4 is then evaluated as -(sum over clipped elements),
 * which removes the cancellation noise of the un-clipped ones, and d_aft (exactly zero) need not be requested.
 */
OFQ_API long long ofq_lsq_bwd_workspace(long long rows, int cols, int nseg);
OFQ_API int ofq_lsq_bwd(const float* dy, long long lddy, const float* x, long long ldx, long long rows, int cols,
                        const float* b4, const float* s_eff, int scale_mode, int period, int nseg,
                        int qlo, int qhi, float* dx, long long lddx, float* workspace, void* stream);
/* Same for codes = Q(act(x) + b4) (ofq_lsq_quant_ex): x is the saved PRE-activation, the quantizer terms use act(x) and
 * dx = dy * inside * act'(x) is the gradient w.r.t. x; d_aft / d_b4 / d_s are unchanged (they live after the activation). */
OFQ_API int ofq_lsq_bwd_act(const float* dy, long long lddy, const float* x, long long ldx, long long rows, int cols,
                            const float* b4, const float* s_eff, int scale_mode, int period, int nseg,
                            int qlo, int qhi, int act, float* dx, long long lddx, float* workspace, void* stream);
/* Same, and the 16-bit operand of the NEXT backward GEMMs written by the same pass (what ofq_grad_prep would make from dx):
 *   out16[r][c] = rn16( dx[r][c] * cs16[c] * rs16[r % rs16_period] * scale4[0] )      (pitch ld16, OFQ_FMT_BF16 / _F16)
 * dx may then be NULL: the gradient of a quantizer input that only feeds a linear layer's backward (the qkx and V
 * quantizers of attention.py:179-206) never exists in fp32 in HBM. scale4 (ofq_scale_from_max / ofq_absmax_scale layout)
 * must bound |dx| <= |dy| up front. Streaming layout only (cols % 4 == 0, segments that are multiples of 128 columns). */
OFQ_API int ofq_lsq_bwd_ex(const float* dy, long long lddy, const float* x, long long ldx, long long rows, int cols,
                           const float* b4, const float* s_eff, int scale_mode, int period, int nseg,
                           int qlo, int qhi, int act, float* dx, long long lddx, void* out16, long long ld16, int fmt16,
                           const float* cs16, const float* rs16, int rs16_period, const float* scale4,
                           float* workspace, void* stream);
/* fp16 range scales (ofq_absmax_scale's out4 layout) from n maxima (e.g. the |output| maximum a GEMM epilogue tracked,
 * ofq_gemm_ex): bound = max(amax) * max|v1| * max|v2| * mult (product != 0), see ofq_absmax_scale. */
OFQ_API int ofq_scale_from_max(const float* amax, int n, const float* v1, int n1, const float* v2, int n2, float mult,
                               int product, float* out4, void* stream);
OFQ_API int ofq_lsq_bwd_finalize(const float* workspace, long long rows, int cols, int scale_mode, int period,
                                 int nseg, float g, float* d_s, float* d_b4, float* d_aft, int zero_sum, void* stream);
/* Same, plus dx_colsum[c] = sum over rows of the dx that ofq_lsq_bwd_ex produced with a fused 16-bit operand (per-row scale
 * mode only; with a fused activation this differs from d_b4 by the activation derivative): colsum(dY) of the next linear
 * layer's backward (its bias gradient and the rank-1 shift term of its dW). */
OFQ_API int ofq_lsq_bwd_finalize_colsum(const float* workspace, long long rows, int cols, int scale_mode, int period,
                                        int nseg, float g, float* d_s, float* d_b4, float* d_aft, int zero_sum,
                                        float* dx_colsum, void* stream);
/* ofq_lsq_bwd_finalize and ofq_lsq_bwd_scale (below) in ONE launch: the reductions of the partial sums and the fp16
 * range scales out4 of the next consumer of dx (autograd of lsq.py:571-602 feeding the dX / dW GEMMs of qlinear.py:69). */
OFQ_API int ofq_lsq_bwd_finalize_scale(const float* workspace, long long rows, int cols, int scale_mode, int period,
                                       int nseg, float g, float* d_s, float* d_b4, float* d_aft, int zero_sum,
                                       const float* v1, int n1, const float* v2, int n2, float mult, int product,
                                       float* out4, void* stream);
/* fp16 range scales (layout of ofq_absmax_scale's out4) for the NEXT consumer of dx, from the per-block max |dx| that
 * ofq_lsq_bwd left in its workspace: bound_1 = max|dx| * max|v1| * mult, bound_2 = max|dx| * max|v2| * mult. */
OFQ_API int ofq_lsq_bwd_scale(const float* workspace, long long rows, int cols, int nseg, const float* v1, int n1,
                              const float* v2, int n2, float mult, int product, float* out4, void* stream);

/* Gradient operand preparation for the bf16 backward GEMMs: one pass over a fp32 gradient x[nb][R][C]:
 *   out_rm[p][b][r][c] = bf16 plane p of ( x * cs[c] )              (row-major, ld = ld_rm; NULL to skip)
 *   out_t [p][b][c][r] = bf16 plane p of ( x * rs[r % rs_period] )  (per-batch transpose, pitch r_pad; NULL to skip)
 *     planes = 1: plane 0 = bf16(v).  planes = 2: plane 0 = hi = bf16(v), plane 1 = lo = bf16(v - hi); feeding both
 *     planes to the GEMM as two outer-K slices gives ~16 mantissa bits (the integer-code operand is exact in bf16).
 *   colsum[c]       = sum_{b,r} x                      (optional; pre-zeroed, atomically accumulated)
 *   rowdot[b][g][r] = sum_{c in group g} x * u[c]      (optional; groups of `group` = 16, 32 or 64 consecutive columns)
 *   out_fmt = OFQ_FMT_F16 (planes must be 1): single fp16 plane (11 significant bits, ~2e-4 gradient error) of
 *     ( x * cs[c] * scale4[0] ) and ( x * rs[r] * scale4[2] ); scale4 comes from ofq_absmax_scale and keeps the
 *     operand inside the fp16 range; the consuming GEMM multiplies its result by scale4[1] / scale4[3].
 *   rm_rowscale != 0: out_rm additionally carries rs[r]: ONE row-major copy ( x * cs[c] * rs[r] * scale4[0] ) then serves
 *     the dX GEMM (K-major, rs undone per output row) and the dW GEMM (MN-major, cs undone per output row).
 */
OFQ_API int ofq_grad_prep(const float* x, int nb, int R, int C, long long ldx, long long bstride_x,
                          const float* cs, const float* rs, int rs_period, int planes, void* out_rm,
                          long long ld_rm, void* out_t, int r_pad, float* colsum, const float* u, int group,
                          float* rowdot, int out_fmt, const float* scale4, int rm_rowscale, void* stream);

/* Range scales for fp16 gradient operands, one read-only pass over x[nb][R][C] (fp32):
 *   bound_c = max |x * cs[c]| * max_i |v1[i]| * mult,   bound_r = max |x * rs[r % rs_period]| * max_i |v2[i]| * mult
 *   out4 = { sc_c, 1/sc_c, sc_r, 1/sc_r } with sc = 2^k such that bound * sc lies in [2^14, 2^15)
 * (cs / rs / v1 / v2 may be NULL = 1). product != 0: a single bound max |x * cs[c] * rs[r]| * max|v1| * max|v2| * mult
 * for an operand that carries both scale vectors (out4[2..3] repeat out4[0..1]). workspace: uint32[ofq_absmax_scale_workspace()], zero-initialised ONCE by the
 * caller and then reusable by every later call on the same stream (the kernel resets its counter). */
OFQ_API long long ofq_absmax_scale_workspace(void);
OFQ_API int ofq_absmax_scale(const float* x, int nb, int R, int C, long long ldx, long long bstride,
                             const float* cs, const float* rs, int rs_period, const float* v1, int n1,
                             const float* v2, int n2, float mult, int product, float* out4, void* workspace, void* stream);

/* int8 codes [nb][R][C] (row stride ld, batch stride bstride) -> bf16, optionally transposed per batch:
 *   transpose = 0: out[b][r][c] (row stride ld_out);   transpose = 1: out[b][c][r] (row pitch ld_out >= R) */
OFQ_API int ofq_codes_to_bf16(const int8_t* codes, int nb, int R, int C, long long ld, long long bstride,
                              void* out, long long ld_out, long long bstride_out, int transpose, void* stream);
/* Same with the 16-bit format chosen by out_fmt (OFQ_FMT_BF16 / OFQ_FMT_F16); codes are exact in both. */
OFQ_API int ofq_codes_to_16(const int8_t* codes, int nb, int R, int C, long long ld, long long bstride,
                            void* out, long long ld_out, long long bstride_out, int transpose, int out_fmt, void* stream);
/* out[row][seg] = sum_{c in segment seg} u[c] * codes[row][c]  (the move_aft shift of one attention operand
 * folded through the other operand's codes; row has nseg segments of cols/nseg columns). */
OFQ_API int ofq_codes_rowdot(const int8_t* codes, long long rows, int cols, long long ld, int nseg,
                             const float* u, float* out, void* stream);
/* int8 codes [nb][R][C] -> int8 transposed [nb][C][r_pad] (V operand of the P.V GEMM). */
OFQ_API int ofq_codes_transpose(const int8_t* codes, int nb, int R, int C, long long ld, long long bstride,
                                int8_t* out, long long ld_out, long long bstride_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Softmax + unsigned LSQ of the attention probabilities (attention.py:96-99 / 212-216,
 * swin_attention_and_mlp.py:201-227):   P = softmax_d( S[z][n][:] + bias[h][n][:] + mask[w][n][:] ),
 * codes = rint( clamp( P / s_eff[n], 0, qhi ) ),  rowsum[z][n] = s_eff[n] * sum_d codes.
 *   S     [nz][N][ld] fp32 (already multiplied by head_dim^-0.5), z = (b*H + h); window index w = b % nW
 *   P     optional fp32 output (saved for backward), same layout as S
 *   codes [nz][N][ldq] int8, columns N..ldq-1 are zero-filled
 */
OFQ_API int ofq_softmax_quant(const float* S, int nz, int N, long long ld, int H, const float* bias,
                              const float* mask, int nW, const float* s_eff, int qhi, float* P, int8_t* codes,
                              long long ldq, float* rowsum, void* stream);

/* Same, plus an optional exact 16-bit copy codes16 [nz][N][ldq] (OFQ_FMT_BF16 / _F16) of the codes written in the same
 * pass (the A operand of the dV GEMM of the backward). Only without bias / mask and with 16-byte aligned rows. */
OFQ_API int ofq_softmax_quant_ex(const float* S, int nz, int N, long long ld, int H, const float* bias,
                                 const float* mask, int nW, const float* s_eff, int qhi, float* P, int8_t* codes,
                                 long long ldq, float* rowsum, void* codes16, int fmt16, void* stream);

/* Backward of softmax + LSQ given dPq = dL/dP_hat [nz][N][ld] and the saved P:
 *   v = P/s;  inside = v <= qhi (v >= 0 always);  dP = dPq * inside
 *   d_s[n] += g_s * sum_{z,d} dPq * (q - inside * v)                  (atomic; d_s pre-zeroed)
 *   dS = alpha * P * (dP - sum_d P*dP)                                 (alpha = head_dim^-0.5 of the logits)
 *   out_a [b][p][h][n][d] = bf16 plane p of ( dS * ca[d] )   ca index (h*N + d) if ca_per_head else d  (pitch ldo)
 *   out_bt[b][p][h][d][n] = bf16 plane p of ( dS * rb[n] )   (transposed, pitch ldo);  planes as in ofq_grad_prep
 *   colsum[z][d]   += sum_n dS              (optional, atomic, pre-zeroed)
 *   out_fmt = OFQ_FMT_F16 (planes = 1): fp16 planes of ( dS * ca[d] * scale4[0] ) and ( dS * rb[n] * scale4[2] )
 *   a_rowscale != 0: out_a = dS * ca[d] * rb[n] (* scale4[0]) and out_bt may be NULL: one copy that the K-major
 *                     dQ-side GEMM and the MN-major dK-side GEMM both read (the other vector is undone per output row)
 *   dS32            = optional fp32 P * (dP - sum_d P*dP) WITHOUT alpha, layout of P: the gradient w.r.t. the
 *                     additive pre-softmax bias (Swin relative-position bias)
 */
OFQ_API int ofq_softmax_quant_bwd(const float* dPq, const float* P, int nz, int N, long long ld, int H,
                                  const float* s_eff, int qhi, float alpha, float g_s, const float* ca,
                                  int ca_per_head, const float* rb, int planes, void* out_a, void* out_bt,
                                  long long ldo, float* colsum, float* d_s, float* dS32, int out_fmt,
                                  const float* scale4, int a_rowscale, void* stream);
/* Same with a partial buffer ds_partial [nz][N] (or NULL): on the vectorised path (single 16-bit output, aligned rows) the
 * per-row scale-gradient terms are stored there and reduced over the slabs by a second tiny launch (d_s is then written,
 * deterministically, instead of accumulated with same-address atomics). */
OFQ_API int ofq_softmax_quant_bwd_ex(const float* dPq, const float* P, int nz, int N, long long ld, int H,
                                     const float* s_eff, int qhi, float alpha, float g_s, const float* ca,
                                     int ca_per_head, const float* rb, int planes, void* out_a, void* out_bt,
                                     long long ldo, float* colsum, float* d_s, float* dS32, int out_fmt,
                                     const float* scale4, int a_rowscale, float* ds_partial, void* stream);

/* K4  W_qk[h] = W_q[h]^T W_k[h] in fp32 (attention.py:190-194) and its backward. wq, wk: [H*hd][C]. */
OFQ_API int ofq_wqk_compose(const float* wq, const float* wk, int H, int hd, int C, float* wqk, void* stream);
OFQ_API int ofq_wqk_compose_bwd(const float* dwqk, const float* wq, const float* wk, int H, int hd, int C,
                                float* dwq, float* dwk, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Step prologue: the weight-side work of a whole model in three launches. The weights, shifts and LSQ step sizes only
 * change at the optimizer step, so every StatsQ code tensor (statsq.py:133-150), every W_qk product (attention.py:190-194)
 * and every effective LSQ step size (lsq.py:593) can be produced once per forward by multi-tensor kernels instead of one
 * small launch per layer (157 launches per DeiT-S step). Job tables are device arrays of these records; a CTA finds its
 * job by binary search over first_block (cumulative CTA count: ceil(rows / 8) per StatsQ job, ceil(n / 256) per scale job).
 */
typedef struct {
    const float* w; const float* aft; const float* bias;       /* aft / bias may be NULL (see ofq_statsq_codes) */
    int8_t* codes; float* colscale; float* inv_colscale; float* colterm;   /* inv_colscale / colterm may be NULL */
    void* codes16;                                                         /* optional exact 16-bit copy of the codes */
    long long ldw; int rows, cols; float n_levels; int first_block;        /* n_levels = 2^(bits-1); codes pitch = cols */
    int f16; int pad;                                                      /* codes16 format: 1 = fp16, 0 = bf16 */
} ofq_statsq_job_t;
typedef struct {
    const float* alpha; float* out; float* out_recip;          /* out_recip may be NULL */
    int n; float g; int first_block; int pad;
} ofq_scale_job_t;
typedef struct { const float* wq; const float* wk; float* wqk; } ofq_wqk_job_t;
OFQ_API int ofq_statsq_codes_multi(const void* table, int n_jobs, int total_blocks, void* stream);
OFQ_API int ofq_lsq_effective_scale_multi(const void* table, int n_jobs, int total_blocks, void* stream);
OFQ_API int ofq_wqk_compose_multi(const void* table, int n_jobs, int H, int hd, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  Fused quantized attention forward, query-key-reparameterised path (attention.py:210-219): ONE kernel computes
 *   S = x_hat k_hat^T * hd^-1/2 (int8 tcgen05 MMA, K = C, accumulator in TMEM) -> softmax (registers / TMEM) -> unsigned LSQ
 *   probability codes -> P_hat V_hat (second int8 MMA from a shared-memory P operand) -> out [B, N, C] fp32.
 * The logits never exist in HBM. Replaces ofq_gemm(I8 scores) + ofq_softmax_quant + ofq_gemm(I8 P.V) bit for bit.
 *   qx   int8 [B, N, C]      codes of the shared quantized input x_hat
 *   qk   int8 [B, N, H, C]   codes of k_hat = Q(W_qk x): key d of head h at ((b*N + d)*H + h)*C
 *   qvT  int8 [B, C, ldv]    codes of v_hat, transposed (keys contiguous, zero padded to ldv, a multiple of 16)
 *   se_x [N], se_k [N*H] (index d*H + h), se_p [N], se_v [C]: effective step sizes; ctS [B*N, H]: code row-dots
 *   sum_c x_aft[c] qk[b,d,h,c] (the logit term of the input shift); v_aft [C]: shift of v_hat; scale = hd^-1/2.
 *   qp   int8 [B*H, N, ldq]  out: probability codes (ldq >= 208, multiple of 16; keys >= N hold 0)
 *   P    optional out: probabilities fp32 [B*H, N, ldS]; qp16 optional out: exact 16-bit copy of the codes (pitch ldq);
 *   rowsum optional out [B*H, N]: se_p[n] * sum_d qp[n, d]; rowstat optional out [B*H, N, 2]: (row maximum of the scaled
 *   logits, sum of exp) - what ofq_qkr_attn_bwd needs to recompute the probabilities bit for bit.
 * Limits: head dimension 64 (C = 64 H), N <= 208 tokens, C <= 384. */
OFQ_API int ofq_qkr_attn_fwd(const int8_t* qx, const int8_t* qk, const int8_t* qvT, long long ldv, int B, int N, int H, int C,
                             const float* se_x, const float* se_k, const float* ctS, float scale, const float* se_p,
                             int qhi, const float* se_v, const float* v_aft, int8_t* qp, long long ldq, float* out,
                             float* P, long long ldS, void* qp16, int fmt16, float* rowsum, float* rowstat, void* stream);
/* Backward of softmax + probability quantizer fused with the GEMMs that feed it (autograd of attention.py:210-216): the
 * logits are recomputed (int8 MMA), dP = dO v_hat^T is a 16-bit MMA, both stay in TMEM; out come the ONE 16-bit operand of the
 * two score-gradient GEMMs, dS16[b,h,n,d] = rn16(dS * se_k[d*H+h] * se_x[n] * sc_out[0]) (pitch ldo, a multiple of 8),
 * colsum[z,d] = sum_n dS[n,d], the step-size gradient d_s[n] of the probability quantizer (ds_part [B*H, N] is scratch) and
 * sc_out[2] = {range scale of dS16, its reciprocal} (an a-priori power-of-two bound).
 *   a16  16-bit [B, N, C]: dO * se_v[c] * se_p[n] * sc_in[0] (ofq_grad_prep), qv16 16-bit [B, N, C]: codes of v_hat,
 *   rowdot [B, H, N]: sum_j dO[b,n,hj] v_aft[hj], rowstat from the forward, sc_in[2] = {scale of a16, reciprocal},
 *   qmax_v = largest |code| of v_hat, g_s = gradient scale of the probability quantizer's step size (lsq.py:582-591). */
OFQ_API int ofq_qkr_attn_bwd(const int8_t* qx, const int8_t* qk, const void* a16, const void* qv16, int fmt16, int B, int N,
                             int H, int C, const float* se_x, const float* se_k, const float* ctS, float scale,
                             const float* se_p, const float* inv_se_p, int qhi, const float* rowstat, const float* rowdot,
                             const float* sc_in, const float* se_v, const float* v_aft, int qmax_v, float g_s, void* dS16,
                             long long ldo, float* colsum, float* ds_part, float* d_s, float* sc_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  CGA: freeze mask (cga.py:450-469) and the masked AdamW step (cga.py:953-1013 + torch.optim.AdamW).
 * ofq_cga_mask writes 1 for frozen, 0 for trainable (the reference's `freeze_idx`).
 * ofq_cga_adamw updates p, exp_avg, exp_avg_sq in place in ONE pass: frozen elements see a zero gradient
 * (moments decay, weight untouched bit-for-bit); bits == 0 disables masking (plain AdamW for the other params).
 * rowstat: float[rows] scratch for the per-row StatsQ scale; kminmax: int32[2] scratch.
 * step_dev: optional device int32 holding the optimizer step t (used instead of `step` so that a captured CUDA
 * graph can be replayed; advance it with ofq_counter_increment once per optimizer step).
 */
OFQ_API int ofq_cga_mask(const float* w, int rows, int cols, int bits, double boundary_range, uint8_t* mask,
                         float* rowstat, int* kminmax, void* stream);
OFQ_API int ofq_cga_adamw(float* p, const float* grad, float* exp_avg, float* exp_avg_sq, long long numel,
                          int rows, int cols, int step, double lr, double beta1, double beta2, double eps,
                          double weight_decay, int bits, double boundary_range, float* rowstat, int* kminmax,
                          uint8_t* mask_out, const int* step_dev, void* stream);
/* dX GEMM of a quantized linear layer with the backward of the layer's INPUT quantizer as its epilogue (north star: STE mask and
 * LSQ step-size gradient fused into the backward GEMM; autograd of reference lsq.py:571-602 on qlinear.py:58-73):
 *   dxhat[m][n] = (sum_k A[m][k] B[n][k]) * rs[m % period] * cs[0]       (as ofq_gemm; never written)
 *   v = (x[m][n] + b4[n]) * rs[m % period]        (rs = 1 / s_eff: the GEMM's row un-scale IS the quantizer's reciprocal step)
 *   dx[m][n]    = dxhat if qlo <= v <= qhi else 0
 *   d_s[i]      = g * sum_{m % period == i} sum_n dxhat * (q - v if inside else q),   q = rint(clamp(v, qlo, qhi))
 *   d_b4[n]     = sum_m dx[m][n]                                          (deterministic: per-warp partials + one finalize)
 *   d_aft[n]    = sum_m dxhat[m][n] = sum_k dy_colsum[k] * colscale[k] * w_codes[k][n]   (optional; exact identity, no pass
 *                 over dxhat: w_codes int8 [K, N] are the layer's weight codes, dy_colsum = colsum(dY), colscale may be NULL)
 * dx_absmax (optional, one float PRE-SET to 0) receives max |dx| (the bound for the fp16 range scale of whatever consumes dx).
 * 16-bit kinds, plain (unbatched) operands, N % 64 == 0. workspace: ofq_gemm_dx_lsq_workspace(M, N) floats. */
OFQ_API long long ofq_gemm_dx_lsq_workspace(int M, int N);
OFQ_API int ofq_gemm_dx_lsq(int kind, const ofq_operand_t* A, const ofq_operand_t* B, int M, int N, int K, const ofq_vec_t* rs,
                            const ofq_vec_t* cs, const float* x, long long ldx, const float* b4, int qlo, int qhi, float g,
                            float* dx, long long lddx, float* d_s, float* d_b4, float* d_aft, const int8_t* w_codes,
                            long long ld_codes, const float* dy_colsum, const float* colscale, float* dx_absmax, float* workspace,
                            void* stream);
/* The finalize pass of the above on raw partials (colpart [nslots][3][cols], rowpart [planes][rows]). */
OFQ_API int ofq_lsq_bwd_finalize_parts(const float* colpart, long long nslots, const float* rowpart, long long rowpart_total,
                                       long long rows, int cols, int period, float g, float* d_s, float* d_b4, float* d_aft,
                                       void* stream);

/* Deployment export (the reference stops at fake-quant floats): StatsQ weight codes at their true width.
 * A b-bit code 2k+1 (k in [-n, n-1], n = 2^(b-1); reference statsq.py:145-147) is stored as u = k + n in b bits, eight codes per
 * b bytes in little-endian bit order; a row of `cols` codes takes ofq_packed_row_bytes(cols, bits) bytes. Round trip is exact. */
OFQ_API long long ofq_packed_row_bytes(int cols, int bits);
OFQ_API int ofq_pack_codes(const int8_t* codes, long long rows, int cols, long long ld, int bits, uint8_t* out, long long ld_out,
                           void* stream);
OFQ_API int ofq_unpack_codes(const uint8_t* in, long long ld_in, long long rows, int cols, int bits, int8_t* codes, long long ld,
                             void* stream);

/* KD losses of the training recipe (reference src/quantization/utils.py:44-77 KLLossSoft / KDLossSoftandHard; train.py:896-910)
 * and their gradients in one pass over the logits:
 *   row_loss[b] = CE(z_hard[b], target[b])                                  (z_hard != NULL; nn.CrossEntropyLoss, class indices)
 *               - sum_k softmax(teacher[b] / T)_k * log_softmax(z_soft[b] / T)_k   (teacher != NULL)
 *   loss        = mean_b row_loss[b]       (summed in a fixed order)
 *   dz_hard / dz_soft = d loss / d z_hard, d loss / d z_soft  [B, K]; when z_hard == z_soft (single-output student) the sum
 *                       of both terms goes to dz_hard.
 * All logits are dense [B, K] fp32 rows; target int64 [B]. */
OFQ_API int ofq_kd_loss(const float* z_hard, const float* z_soft, const float* teacher, const long long* target, int B, int K,
                        float T, float* row_loss, float* loss, float* dz_hard, float* dz_soft, void* stream);

OFQ_API int ofq_counter_increment(int* counter, void* stream);
/* Multi-tensor plain AdamW (all un-masked parameters of one learning-rate group in ONE launch).
 * table: device array of n_entries 48-byte records
 *   { float* p; const float* g; float* m; float* v; int64 numel; float wd (weight decay; 1 - lr*wd is formed in the kernel); int32 first_block; }
 * where first_block is the running sum of ceil(numel / 1024) and total_blocks the final sum. */
OFQ_API int ofq_adamw_multi(const void* table, int n_entries, int total_blocks, int step, double lr, double beta1,
                            double beta2, double eps, const int* step_dev, void* stream);
/* Multi-tensor CGA-masked AdamW: every freeze-masked weight in THREE launches (scratch init, per-row StatsQ statistics and
 * rounding-level range, masked update) instead of three per weight. table: device array of n_entries 72-byte records
 *   { float* p; const float* g; float* m; float* v; float* rowstat (scratch [rows]); int32* kminmax (scratch [2]);
 *     int32 rows, cols; float wd; int32 first_block; int32 first_rowblock; int32 pad }
 * first_block = running sum of ceil(rows*cols / 1024) (total_blocks the final sum), first_rowblock = running sum of
 * ceil(rows / 8) (total_rowblocks the final sum). Same arithmetic per element as ofq_cga_adamw. */
OFQ_API int ofq_cga_adamw_multi(const void* table, int n_entries, int total_blocks, int total_rowblocks, int step, double lr,
                                double beta1, double beta2, double eps, int bits, double boundary_range, const int* step_dev,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host glue around the quantized layers: fp32 LayerNorm (nn.LayerNorm semantics: biased
 * variance, eps inside the square root) of the DeiT / Swin blocks (deit_vision_transformer.py:154-164).
 * mean / rstd [rows] are saved for the backward. workspace: float[ofq_layernorm_bwd_workspace(rows, cols)]. */
OFQ_API int ofq_layernorm_fwd(const float* x, long long rows, int cols, const float* gamma, const float* beta,
                              float eps, float* y, float* mean, float* rstd, void* stream);
/* xsum = x + add and y = LayerNorm(xsum) in one pass (cols <= 512): the residual add in front of a pre-norm LayerNorm
 * (x = x + attn(...); mlp(norm2(x)), deit_vision_transformer.py:156-163). Bit-identical to a separate add + ofq_layernorm_fwd. */
OFQ_API int ofq_layernorm_fwd_add(const float* x, const float* add, long long rows, int cols, const float* gamma,
                                  const float* beta, float eps, float* xsum, float* y, float* mean, float* rstd,
                                  void* stream);
OFQ_API long long ofq_layernorm_bwd_workspace(long long rows, int cols);
OFQ_API int ofq_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                              const float* rstd, long long rows, int cols, float* dx, float* dgamma, float* dbeta,
                              float* workspace, void* stream);
/* Same with the gradient `res` [rows][cols] that arrives over the residual connection around the LayerNorm
 * (x + f(LN(x)), deit_vision_transformer.py:154-164) added in the same pass: dx = res + LayerNorm'(dy). res may be NULL. */
OFQ_API int ofq_layernorm_bwd_res(const float* dy, const float* x, const float* gamma, const float* mean,
                                  const float* rstd, long long rows, int cols, const float* res, float* dx,
                                  float* dgamma, float* dbeta, float* workspace, void* stream);
/* Same, plus blockmax [ofq_layernorm_bwd_nmax(rows, cols)] = per-CTA max |dx| (cols <= 512 only): dx is the residual-stream
 * gradient that the previous proj / fc2 layer's backward consumes as dY, whose fp16 range scale then needs no pass over it
 * (ofq_scale_from_max). */
OFQ_API long long ofq_layernorm_bwd_nmax(long long rows, int cols);
OFQ_API int ofq_layernorm_bwd_max(const float* dy, const float* x, const float* gamma, const float* mean,
                                  const float* rstd, long long rows, int cols, const float* res, float* dx,
                                  float* dgamma, float* dbeta, float* workspace, float* blockmax, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFQ_B200_H */
