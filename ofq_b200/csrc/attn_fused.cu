// Fused quantized attention forward for the query-key-reparameterised path (reference attention.py:210-219):
//
//   S[b,h,n,d] = (x_hat[b,n,:] . k_hat[b,d,h,:]) * hd^-1/2        int8 codes x int8 codes -> exact int32 in TMEM (K = C)
//   P = softmax_d(S)                                               fp32, in registers / TMEM, never in HBM as logits
//   Qp = round(clamp(P / s_p[n], 0, 2^b - 1))                      unsigned LSQ codes (attention.py:215)
//   O[b,n,h*hd+j] = s_p[n] (s_v[hj] sum_d Qp[n,d] Qv[b,d,hj] + v_aft[hj] sum_d Qp[n,d])      second int8 MMA (K = keys)
//
// One persistent CTA per SM walks (batch, head) units. Per unit the key operand k_hat (N x C codes) and the value operand
// (hd x N codes, transposed) are staged ONCE in shared memory by TMA and serve both 128-row query tiles of the unit.
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected lane), warps 4-7 and 8-11 =
// two softmax warpgroups; warpgroup g owns query tile g of every unit, its own TMEM accumulator (S, later reused for O)
// and its own shared-memory P operand, so that the softmax of one tile overlaps the MMAs / loads / epilogue of the other.
// Thread i of a warpgroup owns query row i: tcgen05.ld 32x32b hands it its row of the accumulator, the softmax needs no
// cross-thread reduction at all. Three passes over the row, all on chip: (1) scaled logits + running maximum, logits
// written back to TMEM; (2) exp and row sum, exponentials written back to TMEM; (3) probabilities, codes -> shared-memory
// P operand (128B-swizzled K-major, the layout the UMMA descriptor expects) and -> HBM.
//
// Arithmetic is bit-identical to the three-kernel path it replaces (ofq_gemm I8 -> ofq_softmax_quant vectorised kernel ->
// ofq_gemm I8): the logits are fmaf(acc * se_x[n], cs[d], ct[d]) exactly as the GEMM epilogue forms them, the row sum is
// accumulated in the same order as that kernel's per-lane partial sums + xor-shuffle tree, quotients and codes use the
// same Markstein / guarded-reciprocal sequences. tests/test_gpu_attn_fused.py asserts torch.equal on every output.
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include "ofq_b200.h"
#include "ptx.cuh"
#include "host_util.h"

namespace ofq {
namespace attn {

constexpr int BM = 128;             // query rows per tile (UMMA M)
constexpr int BN = 208;             // keys, padded to a multiple of 16 (UMMA N of the score product)
constexpr int HD = 64;              // head dimension (UMMA N of the P V product)
constexpr int MAXKB = 3;            // 128-byte K blocks of the score product: C <= 384
constexpr int NTHREADS = 384;
constexpr uint32_t K_BLOCK = BN * 128;          // 26 624 B
constexpr uint32_t Q_BLOCK = BM * 128;          // 16 384 B
constexpr uint32_t V_BLOCK = HD * 128;          //  8 192 B
constexpr uint32_t P_BLOCK = BM * 128;          // 16 384 B
constexpr uint32_t OFF_K = 0;
constexpr uint32_t OFF_Q = OFF_K + MAXKB * K_BLOCK;
constexpr uint32_t OFF_V = OFF_Q + MAXKB * Q_BLOCK;
constexpr uint32_t OFF_P = OFF_V + 2 * V_BLOCK;                 // [warpgroup][2 blocks]
constexpr uint32_t OFF_VEC = OFF_P + 2 * 2 * P_BLOCK;           // [warpgroup][parity]{cs[208], ct[208], sev[64], vaft[64]}
constexpr int VEC_FLOATS = 2 * BN + 2 * HD;                      // 544
constexpr uint32_t OFF_BAR = OFF_VEC + 2 * 2 * VEC_FLOATS * 4;
constexpr int NBAR = 14;
constexpr uint32_t SMEM_TOTAL = OFF_BAR + NBAR * 8 + 16;
constexpr size_t DYN_BYTES = SMEM_TOTAL + 1024;
static_assert(DYN_BYTES <= 227 * 1024, "shared memory budget exceeded");
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t S_COL_STRIDE = 256;          // accumulator of warpgroup g at column 256 g (208 columns; 0..63 are reused for O)

struct Params {
    int B, N, H, C, kblocks, units;
    const float* se_x;      // [N]      effective step of the shared input quantizer
    const float* se_k;      // [N * H]  effective step of the qkx quantizer, index d * H + h
    const float* ctS;       // [B * N, H] code row-dots sum_c x_aft[c] qk[b,d,h,c]
    float scale;
    const float* se_p;      // [N]      effective step of the probability quantizer (per query row)
    float qhi;
    const float* se_v;      // [C]
    const float* v_aft;     // [C]
    int8_t* qp; long long ldq;          // out: probability codes [B*H, N, ldq]
    float* out;                          // out: [B, N, C]
    float* P; long long ldS;            // optional out: probabilities [B*H, N, ldS]
    uint16_t* qp16; int f16;            // optional out: exact 16-bit copy of the codes, pitch ldq
    float* rowsum;                       // optional out: s_p[n] * sum_d Qp[n,d], [B*H, N]
    int debug;                           // measurement only (OFQ_ATTN_DEBUG bits): knock out parts of the softmax warps' work
    float2* rowstat;                     // optional out: (row maximum of the scaled logits, sum of exp) [B*H, N]: lets the
                                         // backward recompute the probabilities bit for bit from the codes alone
};

__device__ __forceinline__ float rcp_fast(float s) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(s));
    return r;
}
// exact rint(clamp(p / s, 0, qhi)) with a cheap quotient (same sequence as softmax_quant.cu prob_code)
__device__ __forceinline__ float prob_code(float p, float s, float inv_s, float qhi) {
    float v = __fmul_rn(p, inv_s);
    float r = rintf(v);
    const float dv = fabsf(v - r);
    if (dv > 0.4998f || (dv < 2e-4f && r == qhi)) {
        v = __fdiv_rn(p, s);
        r = rintf(v);
    }
    return fminf(r, qhi);
}
// The conversion pipe (XU: I2F, F2I, FRND, MUFU) runs at a quarter of the FP32 rate and was 81 % busy in the first version
// of this kernel (ncu, profiles/r02_ncu_attn_fwd.md). Everything but the one MUFU.EX2 per probability now stays on the
// FMA / ALU pipes through the 1.5 * 2^23 trick: for |x| < 2^22, x + MAGIC has the integer rint(x) in its low mantissa bits.
constexpr float MAGIC = 12582912.f;                 // 0x4B400000
__device__ __forceinline__ float i2f_small(uint32_t acc) {      // exact (float)(int32_t)acc for |acc| < 2^22
    return __fadd_rn(__uint_as_float(acc + 0x4B400000u), -MAGIC);
}
__device__ __forceinline__ float rint_small(float v) {          // rintf(v) (ties to even) for |v| < 2^22
    return __fadd_rn(__fadd_rn(v, MAGIC), -MAGIC);
}
__device__ __forceinline__ uint32_t pack4_u8(float a, float b, float c, float d) {
    return (uint32_t)(int)a | ((uint32_t)(int)b << 8) | ((uint32_t)(int)c << 16) | ((uint32_t)(int)d << 24);
}
// four small non-negative integer-valued floats -> packed bytes without a float->int conversion: q + MAGIC carries q in its low byte
__device__ __forceinline__ uint32_t pack4_codes(float a, float b, float c, float d) {
    const float2 xy = __fadd2_rn(make_float2(a, b), make_float2(MAGIC, MAGIC)), zw = __fadd2_rn(make_float2(c, d), make_float2(MAGIC, MAGIC));
    return __byte_perm(__byte_perm(__float_as_uint(xy.x), __float_as_uint(xy.y), 0x0040),
                       __byte_perm(__float_as_uint(zw.x), __float_as_uint(zw.y), 0x0040), 0x5410);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// All per-row work walks the accumulator in chunks of 16 columns inside ROLLED loops: the three passes together are ~1 000
// instructions. (Fully unrolled over the 208 columns the kernel was 13 000 instructions = 208 KB of straight-line code that
// eight warps streamed through per tile: instruction fetch, not issue slots, set the pace - 302 us per layer.)
__device__ __forceinline__ void ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// all outstanding TMEM loads of this thread have landed; ties the registers to this point of the volatile-asm order
__device__ __forceinline__ void ld16_done(uint32_t (&r)[16]) {
    tmem_ld_wait();
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    ld16_issue(taddr, r);
    ld16_done(r);
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// pass 1: scaled logits s = fmaf(acc * rs, cs[d], ct[d]) (the GEMM epilogue's formula), running maximum; logits -> TMEM
// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2: two IEEE operations per issue slot) throughout: the kernel is bound by
// instruction issue and dependency latency, not by any memory system.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 i2f_small2(uint32_t a, uint32_t b) {      // exact int32 -> fp32 for |x| < 2^22, no conversion pipe
    return __fadd2_rn(f2(__uint_as_float(a + 0x4B400000u), __uint_as_float(b + 0x4B400000u)), f2(-MAGIC));
}
__device__ __forceinline__ void pass1_math(uint32_t (&r)[16], const float* cs, const float* ct, float rs, float& m) {
    const float2 rs2 = f2(rs);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(cs + j);
        const float4 t4 = *reinterpret_cast<const float4*>(ct + j);
        const float2 s01 = __ffma2_rn(__fmul2_rn(i2f_small2(r[j + 0], r[j + 1]), rs2), f2(c4.x, c4.y), f2(t4.x, t4.y));
        const float2 s23 = __ffma2_rn(__fmul2_rn(i2f_small2(r[j + 2], r[j + 3]), rs2), f2(c4.z, c4.w), f2(t4.z, t4.w));
        m = fmaxf(m, fmaxf(fmaxf(s01.x, s01.y), fmaxf(s23.x, s23.y)));
        r[j + 0] = __float_as_uint(s01.x); r[j + 1] = __float_as_uint(s01.y);
        r[j + 2] = __float_as_uint(s23.x); r[j + 3] = __float_as_uint(s23.y);
    }
}

// pass 2: e = exp(s - m) for 16 columns = four "lanes" of the vectorised softmax kernel, whose lane l owns the columns
// {4l..4l+3, 128+4l..128+4l+3} and adds them left to right: acc.{x,y,z,w} are the running sums of the four lanes
__device__ __forceinline__ void pass2_math(uint32_t (&r)[16], float m, float4& acc) {
    float e[16];
    const float2 nm = f2(-m), l2e = f2(1.4426950408889634f);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        // softmax_exp(s - m) = ex2(rn((s - m) * log2 e)) on two columns at a time
        const float2 t = __fmul2_rn(__fadd2_rn(f2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), nm), l2e);
        asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e[j]) : "f"(t.x));
        asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e[j + 1]) : "f"(t.y));
        r[j] = __float_as_uint(e[j]);
        r[j + 1] = __float_as_uint(e[j + 1]);
    }
    acc.x += e[0];  acc.x += e[1];  acc.x += e[2];  acc.x += e[3];
    acc.y += e[4];  acc.y += e[5];  acc.y += e[6];  acc.y += e[7];
    acc.z += e[8];  acc.z += e[9];  acc.z += e[10]; acc.z += e[11];
    acc.w += e[12]; acc.w += e[13]; acc.w += e[14]; acc.w += e[15];
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct RowOut {
    float* P;            // optional per-row outputs (nullptr: not requested / row outside the matrix)
    uint16_t* qp16;
    int f16;
    int ldS;
};

// pass 3: probabilities and codes of 16 columns starting at key `col0`; codes -> the swizzled K-major P operand in shared
// memory (from where the tensor core reads them and a TMA store writes them to HBM)
__device__ __forceinline__ void pass3_math(uint32_t (&r)[16], int col0, float sum, float rinv, float s, float inv_s, float qhi,
                                           uint8_t* prow_smem, int r7, const RowOut& o, uint32_t& csum) {
    // codes = min(rint(p / s), qhi) with the quotient p * (1/s); the IEEE quotient decides wherever the product lies within
    // 2e-4 of a rounding boundary (same decisions as prob_code / the unfused kernel). The boundary test is one running
    // maximum per element and ONE branch per 16 columns instead of a branch per element.
    float q[16];
    float maxdv = 0.f;
    const float2 rinv2 = f2(rinv), nsum2 = f2(-sum), is2 = f2(inv_s), mg = f2(MAGIC), nmg = f2(-MAGIC), m1 = f2(-1.f);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const float2 e = f2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
        float2 p = __fmul2_rn(e, rinv2);
        p = __ffma2_rn(__ffma2_rn(nsum2, p, e), rinv2, p);   // Markstein correction of the quotient e / sum
        r[j] = __float_as_uint(p.x); r[j + 1] = __float_as_uint(p.y);
        const float2 v = __fmul2_rn(p, is2);
        const float2 rr = __fadd2_rn(__fadd2_rn(v, mg), nmg);            // rint (ties to even) without the conversion pipe
        const float2 dv = __ffma2_rn(rr, m1, v);                         // v - rr (exact)
        maxdv = fmaxf(maxdv, fmaxf(fabsf(dv.x), fabsf(dv.y)));
        q[j] = fminf(rr.x, qhi);
        q[j + 1] = fminf(rr.y, qhi);
    }
    if (maxdv > 0.4998f) {                                   // rare (a few percent of the chunks)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float p = __uint_as_float(r[j]);
            const float v = __fmul_rn(p, inv_s);
            if (fabsf(v - rint_small(v)) > 0.4998f) q[j] = fminf(rintf(__fdiv_rn(p, s)), qhi);
        }
    }
    uint4 pk;
    pk.x = pack4_codes(q[0], q[1], q[2], q[3]);
    pk.y = pack4_codes(q[4], q[5], q[6], q[7]);
    pk.z = pack4_codes(q[8], q[9], q[10], q[11]);
    pk.w = pack4_codes(q[12], q[13], q[14], q[15]);
    csum = __dp4a(pk.x, 0x01010101u, csum);                  // running integer sum of the row's codes
    csum = __dp4a(pk.y, 0x01010101u, csum);
    csum = __dp4a(pk.z, 0x01010101u, csum);
    csum = __dp4a(pk.w, 0x01010101u, csum);
    const int kb = col0 >> 7, c16 = (col0 & 127) >> 4;
    *reinterpret_cast<uint4*>(prow_smem + kb * P_BLOCK + ((c16 ^ r7) << 4)) = pk;
    if (o.qp16) {
        uint4 h0, h1;
        if (o.f16) {
            h0.x = pack_f16x2(q[0], q[1]);   h0.y = pack_f16x2(q[2], q[3]);   h0.z = pack_f16x2(q[4], q[5]);   h0.w = pack_f16x2(q[6], q[7]);
            h1.x = pack_f16x2(q[8], q[9]);   h1.y = pack_f16x2(q[10], q[11]); h1.z = pack_f16x2(q[12], q[13]); h1.w = pack_f16x2(q[14], q[15]);
        } else {
            h0.x = pack_bf16x2(q[0], q[1]);   h0.y = pack_bf16x2(q[2], q[3]);   h0.z = pack_bf16x2(q[4], q[5]);   h0.w = pack_bf16x2(q[6], q[7]);
            h1.x = pack_bf16x2(q[8], q[9]);   h1.y = pack_bf16x2(q[10], q[11]); h1.z = pack_bf16x2(q[12], q[13]); h1.w = pack_bf16x2(q[14], q[15]);
        }
        uint4* hp = reinterpret_cast<uint4*>(o.qp16 + col0);
        hp[0] = h0; hp[1] = h1;
    }
    if (o.P) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (col0 + j < o.ldS)         // the pitch padding (keys N .. ldS-1) receives exact zeros, as in the unfused kernel
                *reinterpret_cast<float4*>(o.P + col0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                        __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
qkr_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;  uint64_t* k_empty = bars + 1;
    uint64_t* q_full = bars + 2;  uint64_t* q_empty = bars + 3;
    uint64_t* v_full = bars + 4;  uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;      // [2]  MMA -> warpgroup: scores of its tile are in TMEM
    uint64_t* s_free = bars + 8;      // [2]  warpgroup -> MMA: accumulator (O read out) may be overwritten
    uint64_t* p_ready = bars + 10;    // [2]  warpgroup -> MMA: P operand written
    uint64_t* o_full = bars + 12;     // [2]  MMA -> warpgroup: P V product complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = ((int)blockIdx.x < p.units) ? (p.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmP);
        mbar_init(k_full, 1); mbar_init(k_empty, 1);
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 4);
            mbar_init(&p_ready[g], 4); mbar_init(&o_full[g], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t kbytes = (uint32_t)p.kblocks;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            for (int i = 0; i < nunits; ++i) {
                const int u = blockIdx.x + i * gridDim.x, b = u / p.H, h = u - b * p.H;
                // keys of (b, h): rows d of k_hat[b, d, h, :]; query tile 0
                mbar_wait(k_empty, (i & 1) ^ 1);
                mbar_arrive_expect_tx(k_full, kbytes * K_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + OFF_K + kb * K_BLOCK, &tmK, k_full, kb * 128, 0, h, b, 0);
                mbar_wait(q_empty, 1);                    // Q uses 2i (tile 0) and 2i + 1 (tile 1): parity (use & 1) ^ 1
                mbar_arrive_expect_tx(q_full, kbytes * Q_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + OFF_Q + kb * Q_BLOCK, &tmQ, q_full, kb * 128, 0, b, 0, 0);
                // values of (b, h), transposed: rows hj, K = keys
                mbar_wait(v_empty, (i & 1) ^ 1);
                mbar_arrive_expect_tx(v_full, 2 * V_BLOCK);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_5d(smem + OFF_V + kb * V_BLOCK, &tmV, v_full, kb * 128, h * HD, b, 0, 0);
                // query tile 1
                mbar_wait(q_empty, 0);
                mbar_arrive_expect_tx(q_full, kbytes * Q_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + OFF_Q + kb * Q_BLOCK, &tmQ, q_full, kb * 128, BM, b, 0, 0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint32_t IDESC_S = umma_idesc(2u, 1u, BM, BN);      // S32 accumulate, signed int8 operands
            const uint32_t IDESC_O = umma_idesc(2u, 1u, BM, HD);
            const uint32_t sK = smem_u32(smem + OFF_K), sQ = smem_u32(smem + OFF_Q), sV = smem_u32(smem + OFF_V);
            auto scores = [&](int g) {
                const uint32_t d = tmem_base + g * S_COL_STRIDE;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    const uint64_t ad = umma_desc_kmajor_sw128(sQ + kb * Q_BLOCK);
                    const uint64_t bd = umma_desc_kmajor_sw128(sK + kb * K_BLOCK);
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk) umma_i8(d, ad + kk * 2, bd + kk * 2, IDESC_S, (kb | kk) != 0);
                }
            };
            auto pv = [&](int g) {
                const uint32_t d = tmem_base + g * S_COL_STRIDE;
                const uint32_t sP = smem_u32(smem + OFF_P + g * 2 * P_BLOCK);
#pragma unroll
                for (uint32_t k = 0; k < 7; ++k) {                     // 7 x 32 = 224 >= 208 keys (columns 208..223 hold zeros)
                    const uint32_t kb = k >> 2, kk = k & 3;
                    umma_i8(d, umma_desc_kmajor_sw128(sP + kb * P_BLOCK) + kk * 2, umma_desc_kmajor_sw128(sV + kb * V_BLOCK) + kk * 2,
                            IDESC_O, k != 0);
                }
            };
            for (int i = 0; i < nunits; ++i) {
                // scores of tile 0
                mbar_wait(k_full, i & 1);
                mbar_wait(q_full, 0);
                mbar_wait(&s_free[0], (i & 1) ^ 1);
                tc_fence_after();
                scores(0);
                tc_commit(q_empty);
                tc_commit(&s_full[0]);
                // P V of tile 1 of the previous unit
                if (i > 0) {
                    mbar_wait(&p_ready[1], (i - 1) & 1);
                    tc_fence_after();
                    pv(1);
                    tc_commit(v_empty);
                    tc_commit(&o_full[1]);
                }
                // scores of tile 1
                mbar_wait(q_full, 1);
                mbar_wait(&s_free[1], (i & 1) ^ 1);
                tc_fence_after();
                scores(1);
                tc_commit(q_empty);
                tc_commit(k_empty);
                tc_commit(&s_full[1]);
                // P V of tile 0
                mbar_wait(v_full, i & 1);
                mbar_wait(&p_ready[0], i & 1);
                tc_fence_after();
                pv(0);
                tc_commit(&o_full[0]);
            }
            if (nunits > 0) {
                mbar_wait(&p_ready[1], (nunits - 1) & 1);
                tc_fence_after();
                pv(1);
                tc_commit(v_empty);
                tc_commit(&o_full[1]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------------------------------ softmax warpgroups
        const int g = (warp - 4) >> 2;               // warpgroup = query tile of every unit
        const int q = warp & 3;                      // TMEM lane quarter of this warp
        const int t = threadIdx.x - 128 - g * 128;   // 0..127: row inside the tile
        const int n = g * BM + t;                    // query row
        const bool row_ok = n < p.N;
        const bool warp_ok = g * BM + q * 32 < p.N;  // warp-uniform: at least one valid row
        const uint32_t trow = tmem_base + g * S_COL_STRIDE + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* pbase = smem + OFF_P + g * 2 * P_BLOCK;
        uint8_t* prow = pbase + t * 128;
        const int r7 = t & 7;
        // keys 208..223 of the P operand are read by the last MMA and never written by data: zero them once
        *reinterpret_cast<uint4*>(prow + P_BLOCK + ((5 ^ r7) << 4)) = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        const float rs = row_ok ? __ldg(p.se_x + n) : 0.f;
        const float s_p = row_ok ? __ldg(p.se_p + n) : 1.f;
        const float inv_s = rcp_fast(s_p);
        float* vecs = reinterpret_cast<float*>(smem + OFF_VEC) + g * 2 * VEC_FLOATS;
        // vector entries this thread stages per unit: keys t and t + 128, and one of se_v / v_aft
        auto load_vecs = [&](int u, float (&v)[5]) {
            const int b = u / p.H, h = u - b * p.H;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int d = t + 128 * k;
                float cs = 0.f, ct = -INFINITY;       // keys beyond N: logit -inf -> probability 0 -> code 0
                if (d < p.N) {
                    cs = __fmul_rn(__ldg(p.se_k + (long long)d * p.H + h), p.scale);
                    ct = __fmul_rn(__ldg(p.ctS + ((long long)b * p.N + d) * p.H + h), cs);
                }
                v[2 * k] = cs; v[2 * k + 1] = ct;
            }
            v[4] = t < HD ? __ldg(p.se_v + h * HD + t) : __ldg(p.v_aft + h * HD + (t - HD));
        };
        float nv[5];
        if (nunits > 0) load_vecs(blockIdx.x, nv);
        for (int i = 0; i < nunits; ++i) {
            const int u = blockIdx.x + i * gridDim.x, b = u / p.H, h = u - b * p.H;
            const long long z = u;                   // (b, h) slab index = b * H + h
            float* vb = vecs + (i & 1) * VEC_FLOATS;
            if (t == 0) tma_store_wait_read<0>();     // last unit's code stores have read the P operand (ordered by the barrier below)
            vb[t] = nv[0]; vb[BN + t] = nv[1];
            if (t + 128 < BN) { vb[t + 128] = nv[2]; vb[BN + t + 128] = nv[3]; }
            vb[2 * BN + t] = nv[4];
            named_bar_sync(1 + g, 128);
            if (i + 1 < nunits) load_vecs(u + gridDim.x, nv);      // next unit's vectors: latency hidden behind this tile
            const float* cs = vb;
            const float* ct = vb + BN;
            const float* sev = vb + 2 * BN;
            const float* vaft = sev + HD;

            mbar_wait(&s_full[g], i & 1);
            tc_fence_after();
            float r_s = 0.f;                         // s_p * sum of the row's codes (rank-1 term of the P V epilogue)
            if (warp_ok) {
                // TMEM loads run one chunk ahead of the arithmetic (two register buffers, loops rolled in pairs)
                constexpr int NCH = BN / 16;                      // 13 chunks of 16 columns
                uint32_t ra[16], rb[16];
                float m = -INFINITY;
                ld16_issue(trow, ra);
#pragma unroll 1
                for (int c = 0; c < NCH; c += 2) {
                    ld16_done(ra);
                    if (c + 1 < NCH) ld16_issue(trow + 16 * (c + 1), rb);
                    pass1_math(ra, cs + 16 * c, ct + 16 * c, rs, m);
                    st16(trow + 16 * c, ra);
                    if (c + 1 < NCH) {
                        ld16_done(rb);
                        if (c + 2 < NCH) ld16_issue(trow + 16 * (c + 2), ra);
                        pass1_math(rb, cs + 16 * (c + 1), ct + 16 * (c + 1), rs, m);
                        st16(trow + 16 * (c + 1), rb);
                    }
                }
                tmem_st_wait();
                // exp + row sum in the summation order of the vectorised softmax kernel: lane sums p[l] of its 32 lanes (columns
                // 4l..4l+3 then 128+4l..128+4l+3), then the xor-shuffle tree p[l] + p[l^16], ... Four lanes per 16-column chunk;
                // chunk order: for c4 = 0..3: lanes 4c4.. (first, second halves), lanes 16+4c4.. (first halves, second: c4 = 0 only)
                float4 a[4];
                {
                    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                    ld16_issue(trow, ra);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        // ra: columns 16 c4 (issued)
                        ld16_done(ra); ld16_issue(trow + 128 + 16 * c4, rb);
                        pass2_math(ra, m, lo); st16(trow + 16 * c4, ra);
                        ld16_done(rb); ld16_issue(trow + 64 + 16 * c4, ra);
                        pass2_math(rb, m, lo); st16(trow + 128 + 16 * c4, rb);
                        ld16_done(ra);
                        if (c4 == 0) ld16_issue(trow + 192, rb); else if (c4 < 3) ld16_issue(trow + 16 * (c4 + 1), rb);
                        pass2_math(ra, m, hi); st16(trow + 64 + 16 * c4, ra);
                        if (c4 == 0) {
                            ld16_done(rb); ld16_issue(trow + 16, ra);
                            pass2_math(rb, m, hi); st16(trow + 192, rb);
                        } else if (c4 < 3) {
                            ld16_done(rb);
#pragma unroll
                            for (int j = 0; j < 16; ++j) ra[j] = rb[j];      // next iteration expects its first chunk in ra
                        }
                        a[c4] = add4(lo, hi);                                // p[l] + p[l + 16]
                        lo = make_float4(0.f, 0.f, 0.f, 0.f); hi = lo;
                    }
                }
                tmem_st_wait();
                const float4 b0 = add4(a[0], a[2]), b1 = add4(a[1], a[3]);           // offset 8
                const float4 c0 = add4(b0, b1);                                      // offset 4
                const float sum = (c0.x + c0.z) + (c0.y + c0.w);                     // offsets 2, 1
                float rinv = rcp_fast(sum);
                rinv = fmaf(rinv, fmaf(-sum, rinv, 1.0f), rinv);
                RowOut ro;
                ro.P = (row_ok && p.P) ? p.P + (z * p.N + n) * p.ldS : nullptr;
                ro.qp16 = (row_ok && p.qp16) ? p.qp16 + (z * p.N + n) * p.ldq : nullptr;
                ro.f16 = p.f16;
                ro.ldS = (int)p.ldS;
                uint32_t csum = 0u;
                ld16_issue(trow, ra);
#pragma unroll 1
                for (int c = 0; c < NCH; c += 2) {
                    ld16_done(ra);
                    if (c + 1 < NCH) ld16_issue(trow + 16 * (c + 1), rb);
                    pass3_math(ra, 16 * c, sum, rinv, s_p, inv_s, p.qhi, prow, r7, ro, csum);
                    if (c + 1 < NCH) {
                        ld16_done(rb);
                        if (c + 2 < NCH) ld16_issue(trow + 16 * (c + 2), ra);
                        pass3_math(rb, 16 * (c + 1), sum, rinv, s_p, inv_s, p.qhi, prow, r7, ro, csum);
                    }
                }
                r_s = __fmul_rn(s_p, (float)csum);
                if (row_ok && p.rowsum) p.rowsum[z * p.N + n] = r_s;
                if (row_ok && p.rowstat) p.rowstat[z * p.N + n] = make_float2(m, sum);
            }
            // P operand (generic-proxy stores) -> visible to the async proxy (tensor core, TMA); all TMEM reads of S are done
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[g]);
            // the codes leave for HBM straight from the P operand: two TMA stores (128 queries x 128 keys each; rows >= N and
            // keys >= ldq are clipped by the tensor map) instead of 13 scattered 16-byte stores per thread
            named_bar_sync(3 + g, 128);
            if (t == 0) {
                tma_store_5d(&tmP, pbase, 0, g * BM, (int)z, 0, 0);
                tma_store_5d(&tmP, pbase + P_BLOCK, 128, g * BM, (int)z, 0, 0);
                tma_store_commit();
            }

            mbar_wait(&o_full[g], i & 1);
            tc_fence_after();
            if (warp_ok) {
#pragma unroll 1
                for (int c = 0; c < HD / 16; ++c) {
                    uint32_t r[16];
                    ld16(trow + 16 * c, r);
                    if (row_ok) {
                        float* orow = p.out + ((long long)b * p.N + n) * p.C + h * HD + 16 * c;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 sv = *reinterpret_cast<const float4*>(sev + 16 * c + j);
                            const float4 va = *reinterpret_cast<const float4*>(vaft + 16 * c + j);
                            const float2 o01 = __ffma2_rn(__fmul2_rn(i2f_small2(r[j + 0], r[j + 1]), f2(s_p)), f2(sv.x, sv.y),
                                                          __fmul2_rn(f2(r_s), f2(va.x, va.y)));
                            const float2 o23 = __ffma2_rn(__fmul2_rn(i2f_small2(r[j + 2], r[j + 3]), f2(s_p)), f2(sv.z, sv.w),
                                                          __fmul2_rn(f2(r_s), f2(va.z, va.w)));
                            *reinterpret_cast<float4*>(orow + j) = make_float4(o01.x, o01.y, o23.x, o23.y);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
        }
        if (t == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Forward, 16 softmax warps (the version the library dispatches). ncu of the 8-warp kernel above: 38 % issue-slot
// utilisation, ~300 cycles of exposed latency per TMEM access (one 16-column load or store per ~100 instructions, two
// warps per scheduler to cover it): the kernel is bound by TMEM round-trip latency, not by issue slots (halving the
// instruction count with packed math did not move it). Here every 128-query tile is shared by TWO warpgroups, which take
// the first / second 8 columns of every 16-column group (the two column sets are the lanes with bit 1 clear / set of the
// vectorised softmax kernel, so its summation tree splits cleanly: each side produces two of the four partial sums c[0..3]
// and only those, the row maximum and the code sums cross through shared memory). Four warps per scheduler, half the
// TMEM accesses per warp.
#ifdef OFQ_ATTN_TRACE
#define ATR_T(t) const long long t = clock64()
#define ATR_ADD(k, t) atr[k] += clock64() - t
#else
#define ATR_T(t)
#define ATR_ADD(k, t)
#endif
constexpr int NTHREADS16 = 576;                 // warp 0 TMA, warp 1 MMA, warps 2..17 softmax
constexpr uint32_t OFF_X16 = OFF_BAR + NBAR * 8 + 16;              // exchange area: [tile][half][128 rows] x {max, c_lo, c_hi, csum}
constexpr size_t DYN_BYTES16 = OFF_X16 + 2 * 2 * BM * 4 * 4 + 1024;
static_assert(DYN_BYTES16 <= 227 * 1024, "shared memory budget exceeded");

__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    tmem_ld_wait();
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// exponentials of 8 columns = two lanes of the vectorised kernel (4 columns each), accumulated left to right
__device__ __forceinline__ void exp8(uint32_t taddr, float m, float2& acc) {
    uint32_t r[8];
    ld8(taddr, r);
    float e[8];
    const float2 nm = f2(-m), l2e = f2(1.4426950408889634f);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const float2 t = __fmul2_rn(__fadd2_rn(f2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), nm), l2e);
        asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e[j]) : "f"(t.x));
        asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e[j + 1]) : "f"(t.y));
        r[j] = __float_as_uint(e[j]);
        r[j + 1] = __float_as_uint(e[j + 1]);
    }
    acc.x += e[0]; acc.x += e[1]; acc.x += e[2]; acc.x += e[3];
    acc.y += e[4]; acc.y += e[5]; acc.y += e[6]; acc.y += e[7];
    st8(taddr, r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }

__global__ void __launch_bounds__(NTHREADS16, 1)
qkr_attn_fwd16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;  uint64_t* k_empty = bars + 1;
    uint64_t* q_full = bars + 2;  uint64_t* q_empty = bars + 3;
    uint64_t* v_full = bars + 4;  uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;
    uint64_t* s_free = bars + 8;
    uint64_t* p_ready = bars + 10;
    uint64_t* o_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = ((int)blockIdx.x < p.units) ? (p.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmP);
        mbar_init(k_full, 1); mbar_init(k_empty, 1);
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 8);
            mbar_init(&p_ready[g], 8); mbar_init(&o_full[g], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t kbytes = (uint32_t)p.kblocks;

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < nunits; ++i) {
                const int u = blockIdx.x + i * gridDim.x, b = u / p.H, h = u - b * p.H;
                mbar_wait(k_empty, (i & 1) ^ 1);
                mbar_arrive_expect_tx(k_full, kbytes * K_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + OFF_K + kb * K_BLOCK, &tmK, k_full, kb * 128, 0, h, b, 0);
                mbar_wait(q_empty, 1);
                mbar_arrive_expect_tx(q_full, kbytes * Q_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + OFF_Q + kb * Q_BLOCK, &tmQ, q_full, kb * 128, 0, b, 0, 0);
                mbar_wait(v_empty, (i & 1) ^ 1);
                mbar_arrive_expect_tx(v_full, 2 * V_BLOCK);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_5d(smem + OFF_V + kb * V_BLOCK, &tmV, v_full, kb * 128, h * HD, b, 0, 0);
                mbar_wait(q_empty, 0);
                mbar_arrive_expect_tx(q_full, kbytes * Q_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + OFF_Q + kb * Q_BLOCK, &tmQ, q_full, kb * 128, BM, b, 0, 0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t IDESC_S = umma_idesc(2u, 1u, BM, BN);
            const uint32_t IDESC_O = umma_idesc(2u, 1u, BM, HD);
            const uint32_t sK = smem_u32(smem + OFF_K), sQ = smem_u32(smem + OFF_Q), sV = smem_u32(smem + OFF_V);
            auto scores = [&](int g) {
                const uint32_t d = tmem_base + g * S_COL_STRIDE;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    const uint64_t ad = umma_desc_kmajor_sw128(sQ + kb * Q_BLOCK);
                    const uint64_t bd = umma_desc_kmajor_sw128(sK + kb * K_BLOCK);
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk) umma_i8(d, ad + kk * 2, bd + kk * 2, IDESC_S, (kb | kk) != 0);
                }
            };
            auto pv = [&](int g) {
                const uint32_t d = tmem_base + g * S_COL_STRIDE;
                const uint32_t sP = smem_u32(smem + OFF_P + g * 2 * P_BLOCK);
#pragma unroll
                for (uint32_t k = 0; k < 7; ++k) {
                    const uint32_t kb = k >> 2, kk = k & 3;
                    umma_i8(d, umma_desc_kmajor_sw128(sP + kb * P_BLOCK) + kk * 2, umma_desc_kmajor_sw128(sV + kb * V_BLOCK) + kk * 2,
                            IDESC_O, k != 0);
                }
            };
            for (int i = 0; i < nunits; ++i) {
                mbar_wait(k_full, i & 1);
                mbar_wait(q_full, 0);
                mbar_wait(&s_free[0], (i & 1) ^ 1);
                tc_fence_after();
                scores(0);
                tc_commit(q_empty);
                tc_commit(&s_full[0]);
                if (i > 0) {
                    mbar_wait(&p_ready[1], (i - 1) & 1);
                    tc_fence_after();
                    pv(1);
                    tc_commit(v_empty);
                    tc_commit(&o_full[1]);
                }
                mbar_wait(q_full, 1);
                mbar_wait(&s_free[1], (i & 1) ^ 1);
                tc_fence_after();
                scores(1);
                tc_commit(q_empty);
                tc_commit(k_empty);
                tc_commit(&s_full[1]);
                mbar_wait(v_full, i & 1);
                mbar_wait(&p_ready[0], i & 1);
                tc_fence_after();
                pv(0);
                tc_commit(&o_full[0]);
            }
            if (nunits > 0) {
                mbar_wait(&p_ready[1], (nunits - 1) & 1);
                tc_fence_after();
                pv(1);
                tc_commit(v_empty);
                tc_commit(&o_full[1]);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------ softmax warps 2..17
        const int cw = warp - 2;
        const int g = cw >> 3;                       // query tile of every unit
        const int hf = (cw >> 2) & 1;                // first / second 8 columns of every 16-column group
        const int q = warp & 3;                      // TMEM lane quarter of this warp (hardware: warp id mod 4)
        const int t = q * 32 + lane;                 // row inside the tile
        const int pt = hf * BM + t;                  // thread index inside the tile's pair of warpgroups (0..255)
        const int n = g * BM + t;
        const bool row_ok = n < p.N;
        const bool warp_ok = g * BM + q * 32 < p.N;
        const uint32_t trow = tmem_base + g * S_COL_STRIDE + (static_cast<uint32_t>(q * 32) << 16) + 8 * hf;
        uint8_t* pbase = smem + OFF_P + g * 2 * P_BLOCK;
        uint8_t* prow = pbase + t * 128;
        const int r7 = t & 7;
        if (hf == 0) *reinterpret_cast<uint4*>(prow + P_BLOCK + ((5 ^ r7) << 4)) = make_uint4(0, 0, 0, 0);   // keys 208..223: zeros
        fence_proxy_async_smem();
        const float rs = row_ok ? __ldg(p.se_x + n) : 0.f;
        const float s_p = row_ok ? __ldg(p.se_p + n) : 1.f;
        const float inv_s = rcp_fast(s_p);
        float* vecs = reinterpret_cast<float*>(smem + OFF_VEC) + g * 2 * VEC_FLOATS;
        float* xch = reinterpret_cast<float*>(smem + OFF_X16) + g * (2 * BM * 4);      // [half][row][4]
        float* xmine = xch + (hf * BM + t) * 4;
        const float* xother = xch + ((hf ^ 1) * BM + t) * 4;
        auto load_vecs = [&](int u, float (&v)[3]) {      // key pt (< 208): cs, ct; one of se_v / v_aft for pt < 128
            const int b = u / p.H, h = u - b * p.H;
            v[0] = 0.f; v[1] = -INFINITY; v[2] = 0.f;
            if (pt < p.N) {
                v[0] = __fmul_rn(__ldg(p.se_k + (long long)pt * p.H + h), p.scale);
                v[1] = __fmul_rn(__ldg(p.ctS + ((long long)b * p.N + pt) * p.H + h), v[0]);
            }
            if (pt < 2 * HD) v[2] = pt < HD ? __ldg(p.se_v + h * HD + pt) : __ldg(p.v_aft + h * HD + (pt - HD));
        };
        float nv[3];
        if (nunits > 0) load_vecs(blockIdx.x, nv);
#ifdef OFQ_ATTN_TRACE
        long long atr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
        for (int i = 0; i < nunits; ++i) {
            const int u = blockIdx.x + i * gridDim.x, b = u / p.H, h = u - b * p.H;
            const long long z = u;
            float* vb = vecs + (i & 1) * VEC_FLOATS;
            if (pt == 0) tma_store_wait_read<0>();    // last unit's code stores have read the P operand (ordered by the barrier below)
            if (pt < BN) { vb[pt] = nv[0]; vb[BN + pt] = nv[1]; }
            if (pt < 2 * HD) vb[2 * BN + pt] = nv[2];
            named_bar_sync(1 + g, 256);
            if (i + 1 < nunits) load_vecs(u + gridDim.x, nv);
            const float* cs = vb + 8 * hf;
            const float* ct = vb + BN + 8 * hf;
            const float* sev = vb + 2 * BN;
            const float* vaft = sev + HD;

            ATR_T(t0);
            mbar_wait(&s_full[g], i & 1);
            tc_fence_after();
            ATR_ADD(0, t0);
            ATR_T(t1);
            float m = -INFINITY;
            if (warp_ok && !(p.debug & 1)) {
                // ---- pass 1: scaled logits of this side's 104 columns, maximum; logits -> TMEM
#pragma unroll 1
                for (int c = 0; c < BN / 16; ++c) {
                    uint32_t r[8];
                    ld8(trow + 16 * c, r);
                    const float4 c0 = *reinterpret_cast<const float4*>(cs + 16 * c), c1 = *reinterpret_cast<const float4*>(cs + 16 * c + 4);
                    const float4 t0 = *reinterpret_cast<const float4*>(ct + 16 * c), t1 = *reinterpret_cast<const float4*>(ct + 16 * c + 4);
                    const float2 rs2 = f2(rs);
                    const float2 s01 = __ffma2_rn(__fmul2_rn(i2f_small2(r[0], r[1]), rs2), f2(c0.x, c0.y), f2(t0.x, t0.y));
                    const float2 s23 = __ffma2_rn(__fmul2_rn(i2f_small2(r[2], r[3]), rs2), f2(c0.z, c0.w), f2(t0.z, t0.w));
                    const float2 s45 = __ffma2_rn(__fmul2_rn(i2f_small2(r[4], r[5]), rs2), f2(c1.x, c1.y), f2(t1.x, t1.y));
                    const float2 s67 = __ffma2_rn(__fmul2_rn(i2f_small2(r[6], r[7]), rs2), f2(c1.z, c1.w), f2(t1.z, t1.w));
                    m = fmaxf(m, fmaxf(fmaxf(fmaxf(s01.x, s01.y), fmaxf(s23.x, s23.y)), fmaxf(fmaxf(s45.x, s45.y), fmaxf(s67.x, s67.y))));
                    r[0] = __float_as_uint(s01.x); r[1] = __float_as_uint(s01.y); r[2] = __float_as_uint(s23.x); r[3] = __float_as_uint(s23.y);
                    r[4] = __float_as_uint(s45.x); r[5] = __float_as_uint(s45.y); r[6] = __float_as_uint(s67.x); r[7] = __float_as_uint(s67.y);
                    st8(trow + 16 * c, r);
                }
                tmem_st_wait();
            }
            ATR_ADD(1, t1);
            ATR_T(t2);
            xmine[0] = m;
            named_bar_sync(1 + g, 256);
            m = fmaxf(m, xother[0]);
            ATR_ADD(2, t2);
            ATR_T(t3);
            float2 cc = f2(0.f);
            if (warp_ok && !(p.debug & 2)) {
                // ---- pass 2: exponentials; this side owns the lanes 4k + 2 hf, 4k + 2 hf + 1 (k = 0..7) of the vectorised kernel:
                //      lane sums P[k] (first half: columns 16 k + 8 hf .., second half: 128 + 16 k + 8 hf .. while < 208), then its
                //      xor tree: offset 16 pairs k with k + 4, offset 8 k with k + 2, offset 4 k with k + 1 -> c[2 hf], c[2 hf + 1]
                float2 P8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    P8[k] = f2(0.f);
                    exp8(trow + 16 * k, m, P8[k]);
                    if (128 + 16 * k < BN) exp8(trow + 128 + 16 * k, m, P8[k]);
                }
                tmem_st_wait();
                const float2 a0 = add2(P8[0], P8[4]), a1 = add2(P8[1], P8[5]), a2 = add2(P8[2], P8[6]), a3 = add2(P8[3], P8[7]);
                cc = add2(add2(a0, a2), add2(a1, a3));
            }
            ATR_ADD(3, t3);
            ATR_T(t4);
            xmine[1] = cc.x; xmine[2] = cc.y;
            named_bar_sync(1 + g, 256);
            ATR_ADD(4, t4);
            ATR_T(t5);
            const float2 co = f2(xother[1], xother[2]);
            const float2 clo = hf ? co : cc, chi = hf ? cc : co;            // (c[0], c[1]) and (c[2], c[3])
            const float sum = (clo.x + chi.x) + (clo.y + chi.y);          // offsets 2, 1 of the tree
            float rinv = rcp_fast(sum);
            rinv = fmaf(rinv, fmaf(-sum, rinv, 1.0f), rinv);
            uint32_t csum = 0u;
            if (warp_ok && !(p.debug & 4)) {
                // ---- pass 3: probabilities and codes of this side's columns
                float* Prow = (row_ok && p.P) ? p.P + (z * p.N + n) * p.ldS : nullptr;
                uint16_t* hrow = (row_ok && p.qp16) ? p.qp16 + (z * p.N + n) * p.ldq : nullptr;
#pragma unroll 1
                for (int c = 0; c < BN / 16; ++c) {
                    uint32_t r[8];
                    ld8(trow + 16 * c, r);
                    float q8[8];
                    float maxdv = 0.f;
                    const float2 rinv2 = f2(rinv), nsum2 = f2(-sum), is2 = f2(inv_s), mg = f2(MAGIC), nmg = f2(-MAGIC), m1 = f2(-1.f);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const float2 e = f2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                        float2 pp = __fmul2_rn(e, rinv2);
                        pp = __ffma2_rn(__ffma2_rn(nsum2, pp, e), rinv2, pp);
                        r[j] = __float_as_uint(pp.x); r[j + 1] = __float_as_uint(pp.y);
                        const float2 v = __fmul2_rn(pp, is2);
                        const float2 rr = __fadd2_rn(__fadd2_rn(v, mg), nmg);
                        const float2 dv = __ffma2_rn(rr, m1, v);
                        maxdv = fmaxf(maxdv, fmaxf(fabsf(dv.x), fabsf(dv.y)));
                        q8[j] = fminf(rr.x, p.qhi);
                        q8[j + 1] = fminf(rr.y, p.qhi);
                    }
                    if (maxdv > 0.4998f) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float pp = __uint_as_float(r[j]);
                            const float v = __fmul_rn(pp, inv_s);
                            if (fabsf(v - rint_small(v)) > 0.4998f) q8[j] = fminf(rintf(__fdiv_rn(pp, s_p)), p.qhi);
                        }
                    }
                    uint2 pk;
                    pk.x = pack4_codes(q8[0], q8[1], q8[2], q8[3]);
                    pk.y = pack4_codes(q8[4], q8[5], q8[6], q8[7]);
                    csum = __dp4a(pk.x, 0x01010101u, csum);
                    csum = __dp4a(pk.y, 0x01010101u, csum);
                    const int col0 = 16 * c;
                    const int kb = col0 >> 7, c16 = (col0 & 127) >> 4;
                    *reinterpret_cast<uint2*>(prow + kb * P_BLOCK + ((c16 ^ r7) << 4) + 8 * hf) = pk;
                    if (hrow) {
                        uint4 h4;
                        if (p.f16) { h4.x = pack_f16x2(q8[0], q8[1]); h4.y = pack_f16x2(q8[2], q8[3]); h4.z = pack_f16x2(q8[4], q8[5]); h4.w = pack_f16x2(q8[6], q8[7]); }
                        else { h4.x = pack_bf16x2(q8[0], q8[1]); h4.y = pack_bf16x2(q8[2], q8[3]); h4.z = pack_bf16x2(q8[4], q8[5]); h4.w = pack_bf16x2(q8[6], q8[7]); }
                        *reinterpret_cast<uint4*>(hrow + col0 + 8 * hf) = h4;
                    }
                    if (Prow) {
                        const int d0 = col0 + 8 * hf;
                        if (d0 < (int)p.ldS) *reinterpret_cast<float4*>(Prow + d0) = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                        if (d0 + 4 < (int)p.ldS) *reinterpret_cast<float4*>(Prow + d0 + 4) = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                    }
                }
            }
            ATR_ADD(5, t5);
            ATR_T(t6);
            xmine[3] = __uint_as_float(csum);
            // P operand (generic-proxy stores) -> visible to the async proxy (tensor core, TMA); all TMEM reads of S are done
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[g]);
            named_bar_sync(1 + g, 256);
            if (pt == 0 && !(p.debug & 32)) {
                tma_store_5d(&tmP, pbase, 0, g * BM, (int)z, 0, 0);
                tma_store_5d(&tmP, pbase + P_BLOCK, 128, g * BM, (int)z, 0, 0);
                tma_store_commit();
            }
            const float r_s = __fmul_rn(s_p, (float)(csum + __float_as_uint(xother[3])));
            if (hf == 0 && row_ok) {
                if (p.rowsum) p.rowsum[z * p.N + n] = r_s;
                if (p.rowstat) p.rowstat[z * p.N + n] = make_float2(m, sum);
            }

            ATR_ADD(6, t6);
            ATR_T(t7);
            mbar_wait(&o_full[g], i & 1);
            tc_fence_after();
            ATR_ADD(7, t7);
            ATR_T(t8);
            if (warp_ok && !(p.debug & 8)) {
                // ---- P V epilogue: this side takes 32 of the head's 64 channels
                const uint32_t orow_t = tmem_base + g * S_COL_STRIDE + (static_cast<uint32_t>(q * 32) << 16) + 32 * hf;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[8];
                    ld8(orow_t + 8 * c, r);
                    if (row_ok) {
                        const int j0 = 32 * hf + 8 * c;
                        float* orow = p.out + ((long long)b * p.N + n) * p.C + h * HD + j0;
#pragma unroll
                        for (int j = 0; j < 8; j += 4) {
                            const float4 sv = *reinterpret_cast<const float4*>(sev + j0 + j);
                            const float4 va = *reinterpret_cast<const float4*>(vaft + j0 + j);
                            const float2 o01 = __ffma2_rn(__fmul2_rn(i2f_small2(r[j + 0], r[j + 1]), f2(s_p)), f2(sv.x, sv.y),
                                                          __fmul2_rn(f2(r_s), f2(va.x, va.y)));
                            const float2 o23 = __ffma2_rn(__fmul2_rn(i2f_small2(r[j + 2], r[j + 3]), f2(s_p)), f2(sv.z, sv.w),
                                                          __fmul2_rn(f2(r_s), f2(va.z, va.w)));
                            *reinterpret_cast<float4*>(orow + j) = make_float4(o01.x, o01.y, o23.x, o23.y);
                        }
                    }
                }
            }
            ATR_ADD(8, t8);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
        }
#ifdef OFQ_ATTN_TRACE
        if (blockIdx.x == 0 && lane == 0 && (cw == 0 || cw == 4 || cw == 8))
            printf("attn fwd16 warp %d (tile %d half %d): units %d | wait S %lld pass1 %lld bar %lld pass2 %lld bar %lld pass3 %lld fence+bar %lld wait O %lld epilogue %lld\n",
                   warp, g, hf, nunits, atr[0], atr[1], atr[2], atr[3], atr[4], atr[5], atr[6], atr[7], atr[8]);
#endif
        if (pt == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================================================
// Backward of softmax + probability quantizer, fused with the two GEMMs that feed it (autograd of attention.py:210-216):
//
//   S  = x_hat k_hat^T * hd^-1/2                      recomputed: int8 MMA into TMEM, as in the forward
//   dP = dO v_hat^T                                    fp16 MMA into TMEM (A16 = dO * se_v[c] * se_p[n] * sc, codes of v)
//   P  = softmax(S) recomputed bit for bit from the forward's row statistics (max, sum of exp)
//   straight-through mask of the probability quantizer, ds_p partial sums, dS = alpha P (dP_masked - <P, dP_masked>)
//   -> dS16[b,h,n,d] = fp16(dS * se_k[h,d] * se_x[n] * sc2)   the ONE operand both score-gradient GEMMs read
//   -> colsum[z,d] = sum_n dS[n,d] (rank-1 term of d k_hat), ds_part[z,n] (step-size gradient of the quantizer)
//
// Neither the logits S, nor the probabilities P, nor dP ever exist in HBM (the three-kernel path wrote dP in fp32, and read
// it and P back: 366 MB per layer at DeiT-S / batch 128). One accumulator pair (S at TMEM column 0, dP at 256) per CTA:
// all eight compute warps work on the same 128-query tile, warpgroup 0 on keys 0..127, warpgroup 1 on keys 128..207; the
// two row-wise scalars a thread needs from the other half (<P, dP> and the ds_p term) cross through shared memory.
constexpr uint32_t B_OFF_K = 0;
constexpr uint32_t B_OFF_Q = B_OFF_K + MAXKB * K_BLOCK;
constexpr uint32_t B_OFF_A = B_OFF_Q + MAXKB * Q_BLOCK;            // A16 tile: 128 queries x 64 channels of the head (16-bit)
constexpr uint32_t B_OFF_V = B_OFF_A + Q_BLOCK;                    // v codes (16-bit): 208 keys x 64 channels
constexpr uint32_t B_OFF_VEC = B_OFF_V + K_BLOCK;                  // [parity]{cs[208], ct[208], ca[208]}
constexpr int B_VEC_FLOATS = 3 * BN;
constexpr uint32_t B_OFF_RED = B_OFF_VEC + 2 * B_VEC_FLOATS * 4;   // [4 column parts][2]{dot, dsp}[128 rows]
constexpr uint32_t B_OFF_COL = B_OFF_RED + 4 * 2 * BM * 4;         // [4 lane quarters][208 keys] column sums of dS
constexpr uint32_t B_OFF_BAR = B_OFF_COL + 4 * BN * 4;
constexpr int B_NBAR = 6;
constexpr size_t B_DYN_BYTES = B_OFF_BAR + B_NBAR * 8 + 16 + 1024;
static_assert(B_DYN_BYTES <= 227 * 1024, "shared memory budget exceeded");
constexpr uint32_t DP_COL = 256;

struct BwdParams {
    int B, N, H, C, kblocks, units;
    const float* se_x; const float* se_k; const float* ctS; float scale;
    const float* se_p; const float* inv_se_p; float qhi;
    const float2* rowstat;      // [B*H, N] (max, sum) from the forward
    const float* rowdot;        // [B, H, N] sum_j dO[b,n,hj] v_aft[hj]
    const float* sc_in;         // [2] range scale of A16 and its reciprocal
    const float* sc_out;        // [2] range scale of dS16 (ofq attn_bwd_scale_kernel) and its reciprocal
    uint16_t* dS16; long long ldo; int f16;
    float* colsum;              // [B*H, N]
    float* ds_part;             // [B*H, N]
};

template <bool F16>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }

template <bool F16>
__global__ void __launch_bounds__(NTHREADS16, 1)
qkr_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmV, const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
    uint64_t* kv_full = bars + 0;  uint64_t* kv_empty = bars + 1;     // keys + values of a unit
    uint64_t* qa_full = bars + 2;  uint64_t* qa_empty = bars + 3;     // queries + gradient operand of a tile
    uint64_t* acc_full = bars + 4; uint64_t* acc_free = bars + 5;     // S and dP of a tile in TMEM / read out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = ((int)blockIdx.x < p.units) ? (p.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmV);
        mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
        mbar_init(qa_full, 1); mbar_init(qa_empty, 1);
        mbar_init(acc_full, 1); mbar_init(acc_free, 16);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t kbytes = (uint32_t)p.kblocks;

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < nunits; ++i) {
                const int u = blockIdx.x + i * gridDim.x, b = u / p.H, h = u - b * p.H;
                mbar_wait(kv_empty, (i & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full, kbytes * K_BLOCK + K_BLOCK);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_5d(smem + B_OFF_K + kb * K_BLOCK, &tmK, kv_full, kb * 128, 0, h, b, 0);
                tma_load_5d(smem + B_OFF_V, &tmV, kv_full, h * HD, 0, b, 0, 0);
                for (int tile = 0; tile < 2; ++tile) {
                    const int use = 2 * i + tile;
                    mbar_wait(qa_empty, (use & 1) ^ 1);
                    mbar_arrive_expect_tx(qa_full, kbytes * Q_BLOCK + Q_BLOCK);
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tma_load_5d(smem + B_OFF_Q + kb * Q_BLOCK, &tmQ, qa_full, kb * 128, tile * BM, b, 0, 0);
                    tma_load_5d(smem + B_OFF_A, &tmA, qa_full, h * HD, tile * BM, b, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t IDESC_S = umma_idesc(2u, 1u, BM, BN);
            const uint32_t IDESC_D = umma_idesc(1u, F16 ? 0u : 1u, BM, BN);      // F32 accumulate, fp16 / bf16 operands
            const uint32_t sK = smem_u32(smem + B_OFF_K), sQ = smem_u32(smem + B_OFF_Q);
            const uint64_t adD = umma_desc_kmajor_sw128(smem_u32(smem + B_OFF_A));
            const uint64_t bdD = umma_desc_kmajor_sw128(smem_u32(smem + B_OFF_V));
            for (int i = 0; i < nunits; ++i) {
                for (int tile = 0; tile < 2; ++tile) {
                    const int use = 2 * i + tile;
                    if (tile == 0) mbar_wait(kv_full, i & 1);
                    mbar_wait(qa_full, use & 1);
                    mbar_wait(acc_free, (use & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        const uint64_t ad = umma_desc_kmajor_sw128(sQ + kb * Q_BLOCK);
                        const uint64_t bd = umma_desc_kmajor_sw128(sK + kb * K_BLOCK);
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) umma_i8(tmem_base, ad + kk * 2, bd + kk * 2, IDESC_S, (kb | kk) != 0);
                    }
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk) umma_f16(tmem_base + DP_COL, adD + kk * 2, bdD + kk * 2, IDESC_D, kk != 0);
                    tc_commit(qa_empty);
                    if (tile == 1) tc_commit(kv_empty);
                    tc_commit(acc_full);
                }
            }
        }
    } else {
        // 16 compute warps on one tile: part = 0..3 takes the 8-key groups part, part + 4, part + 8, ... (26 groups of the 208 keys)
        const int cw = warp - 2;
        const int part = cw >> 2;
        const int q = warp & 3;
        const int t = q * 32 + lane;                 // row inside the tile
        const int tid = threadIdx.x - 64;             // 0..511
        const uint32_t trS = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t trD = trS + DP_COL;
        float* vecs = reinterpret_cast<float*>(smem + B_OFF_VEC);
        float* red = reinterpret_cast<float*>(smem + B_OFF_RED);          // [part][{dot, dsp}][row]
        float* colpart = reinterpret_cast<float*>(smem + B_OFF_COL);      // [quarter][key]
        const float inv_sc_in = __ldg(p.sc_in + 1);
        const float sc_out = __ldg(p.sc_out);
        // key (inside an 8-key group) this lane ends up with after the halving reduction over the warp's 32 rows
        const int mycol = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        auto load_vecs = [&](int u, float (&v)[3]) {
            const int b = u / p.H, h = u - b * p.H;
            v[0] = 0.f; v[1] = -INFINITY; v[2] = 0.f;
            if (tid < BN && tid < p.N) {
                const float ca = __ldg(p.se_k + (long long)tid * p.H + h);
                v[2] = ca;
                v[0] = __fmul_rn(ca, p.scale);
                v[1] = __fmul_rn(__ldg(p.ctS + ((long long)b * p.N + tid) * p.H + h), v[0]);
            }
        };
        float nv[3];
        if (nunits > 0) load_vecs(blockIdx.x, nv);
        for (int k = tid; k < 4 * BN; k += 512) colpart[k] = 0.f;      // (ordered before its first use by the barrier below)
        for (int i = 0; i < nunits; ++i) {
            const int u = blockIdx.x + i * gridDim.x, b = u / p.H, h = u - b * p.H;
            const long long z = u;
            float* vb = vecs + (i & 1) * B_VEC_FLOATS;
            if (tid < BN) { vb[tid] = nv[0]; vb[BN + tid] = nv[1]; vb[2 * BN + tid] = nv[2]; }
            named_bar_sync(1, 512);
            if (i + 1 < nunits) load_vecs(u + gridDim.x, nv);
            const float* cs = vb;
            const float* ct = vb + BN;
            const float* ca = vb + 2 * BN;
            for (int tile = 0; tile < 2; ++tile) {
                const int use = 2 * i + tile;
                const int n = tile * BM + t;
                const bool row_ok = n < p.N;
                const bool warp_ok = tile * BM + q * 32 < p.N;
                float rs = 0.f, s_p = 1.f, inv_sep = 0.f, m = 0.f, sum = 1.f, rdot = 0.f, sca_row = 0.f;
                if (row_ok) {
                    rs = __ldg(p.se_x + n); s_p = __ldg(p.se_p + n); inv_sep = __ldg(p.inv_se_p + n);
                    const float2 st = __ldg(p.rowstat + z * p.N + n);
                    m = st.x; sum = st.y;
                    rdot = __ldg(p.rowdot + ((long long)b * p.H + h) * p.N + n);
                    sca_row = __fmul_rn(sc_out, rs);
                }
                const float inv_s = rcp_fast(s_p);
                float rinv = rcp_fast(sum);
                rinv = fmaf(rinv, fmaf(-sum, rinv, 1.0f), rinv);
                mbar_wait(acc_full, use & 1);
                tc_fence_after();
                float dot = 0.f, dsp = 0.f;
                if (warp_ok) {
                    // ---- pass 1: probabilities, codes, straight-through mask; P -> S columns, masked dP -> dP columns
#pragma unroll 1
                    for (int gi = part; gi < BN / 8; gi += 4) {
                        const int d0 = 8 * gi;
                        uint32_t rS[8], rD[8];
                        ld8(trS + d0, rS);
                        ld8(trD + d0, rD);
                        float pr[8], vq[8];
                        float flag = 0.f;
                        const float4 c0 = *reinterpret_cast<const float4*>(cs + d0), c1 = *reinterpret_cast<const float4*>(cs + d0 + 4);
                        const float4 t0 = *reinterpret_cast<const float4*>(ct + d0), t1 = *reinterpret_cast<const float4*>(ct + d0 + 4);
                        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                        const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                        for (int k = 0; k < 8; k += 2) {
                            const float2 sl = __ffma2_rn(__fmul2_rn(i2f_small2(rS[k], rS[k + 1]), f2(rs)), f2(cc[k], cc[k + 1]), f2(tt[k], tt[k + 1]));
                            const float2 tx = __fmul2_rn(__fadd2_rn(sl, f2(-m)), f2(1.4426950408889634f));
                            float2 e;
                            asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e.x) : "f"(tx.x));
                            asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e.y) : "f"(tx.y));
                            float2 pp = __fmul2_rn(e, f2(rinv));
                            pp = __ffma2_rn(__ffma2_rn(f2(-sum), pp, e), f2(rinv), pp);
                            pr[k] = pp.x; pr[k + 1] = pp.y;
                            const float2 v = __fmul2_rn(pp, f2(inv_s));
                            vq[k] = v.x; vq[k + 1] = v.y;
                            const float2 rr = __fadd2_rn(__fadd2_rn(v, f2(MAGIC)), f2(-MAGIC));
                            const float2 dv = __ffma2_rn(rr, f2(-1.f), v);
                            // rounding boundary, or the clamp bound itself (it decides the straight-through mask)
                            flag = fmaxf(flag, (fabsf(dv.x) > 0.4998f || (fabsf(dv.x) < 2e-4f && rr.x == p.qhi)) ? 1.f : 0.f);
                            flag = fmaxf(flag, (fabsf(dv.y) > 0.4998f || (fabsf(dv.y) < 2e-4f && rr.y == p.qhi)) ? 1.f : 0.f);
                        }
                        if (flag != 0.f) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float rr = rint_small(vq[j]);
                                const float dv = fabsf(vq[j] - rr);
                                if (dv > 0.4998f || (dv < 2e-4f && rr == p.qhi)) vq[j] = __fdiv_rn(pr[j], s_p);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = vq[j];
                            const float qq = fminf(rint_small(v), p.qhi);
                            const bool inside = v <= p.qhi;                   // v >= 0 always holds for probabilities
                            const float gq = fmaf(__fmul_rn(__uint_as_float(rD[j]), inv_sep), inv_sc_in, rdot);
                            dsp = fmaf(gq, inside ? (qq - v) : qq, dsp);
                            const float dp = inside ? gq : 0.f;
                            dot = fmaf(pr[j], dp, dot);
                            rS[j] = __float_as_uint(pr[j]);
                            rD[j] = __float_as_uint(dp);
                        }
                        st8(trS + d0, rS);
                        st8(trD + d0, rD);
                    }
                    tmem_st_wait();
                }
                // (write-after-read on `red` across tiles: the next tile's writes come after its acc_full, which the MMA warp
                // can only signal once ALL sixteen warps have arrived on acc_free, i.e. after every read below - ordered
                // through the mbarrier chain, which compute-sanitizer's racecheck does not follow and reports as a hazard)
                red[(part * 2 + 0) * BM + t] = dot;
                red[(part * 2 + 1) * BM + t] = dsp;
                named_bar_sync(2, 512);
                dot = (red[(0 * 2 + 0) * BM + t] + red[(1 * 2 + 0) * BM + t]) + (red[(2 * 2 + 0) * BM + t] + red[(3 * 2 + 0) * BM + t]);
                if (part == 0 && row_ok)
                    p.ds_part[z * p.N + n] = (red[(0 * 2 + 1) * BM + t] + red[(1 * 2 + 1) * BM + t]) + (red[(2 * 2 + 1) * BM + t] + red[(3 * 2 + 1) * BM + t]);
                if (warp_ok) {
                    // ---- pass 2: dS = alpha P (dP - <P, dP>); scaled 16-bit copy -> HBM; column sums over the warp's rows
                    uint16_t* orow = p.dS16 + (z * p.N + n) * p.ldo;
#pragma unroll 1
                    for (int gi = part; gi < BN / 8; gi += 4) {
                        const int d0 = 8 * gi;
                        uint32_t rS[8], rD[8];
                        ld8(trS + d0, rS);
                        ld8(trD + d0, rD);
                        float ds[8];
                        uint32_t pk[4];
                        const float4 a0 = *reinterpret_cast<const float4*>(ca + d0), a1 = *reinterpret_cast<const float4*>(ca + d0 + 4);
                        const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int k = 0; k < 8; k += 2) {
                            const float r0 = __uint_as_float(rS[k]) * (__uint_as_float(rD[k]) - dot);
                            const float r1 = __uint_as_float(rS[k + 1]) * (__uint_as_float(rD[k + 1]) - dot);
                            ds[k] = row_ok ? p.scale * r0 : 0.f;
                            ds[k + 1] = row_ok ? p.scale * r1 : 0.f;
                            pk[k / 2] = pack16<F16>(ds[k] * aa[k] * sca_row, ds[k + 1] * aa[k + 1] * sca_row);
                        }
                        if (row_ok && d0 < p.ldo) *reinterpret_cast<uint4*>(orow + d0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        // sum of each of the 8 keys over the warp's 32 rows: recursive halving (9 shuffles instead of 40)
                        float w4[4], w2[2], w1;
                        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float send = h16 ? ds[k] : ds[k + 4], keep = h16 ? ds[k + 4] : ds[k];
                            w4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float send = h8 ? w4[k] : w4[k + 2], keep = h8 ? w4[k + 2] : w4[k];
                            w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                        {
                            const float send = h4 ? w2[0] : w2[1], keep = h4 ? w2[1] : w2[0];
                            w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                        w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
                        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                        if ((lane & 3) == 0) colpart[q * BN + d0 + mycol] += w1;      // one lane per (quarter, key): no race
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_free);
            }
            // column sums of the unit: the four lane quarters in a fixed order (quarters without rows in tile 1 added nothing)
            named_bar_sync(3, 512);
            if (tid < p.N) {
                float a = colpart[tid];
                colpart[tid] = 0.f;                          // each reader clears exactly what it read, for the next unit
#pragma unroll
                for (int qq = 1; qq < 4; ++qq) { a += colpart[qq * BN + tid]; colpart[qq * BN + tid] = 0.f; }
                p.colsum[z * p.N + tid] = a;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// fp16 range scale of dS16 from an a-priori bound (the maximum of dP is not known before the fused kernel has run):
//   |dO| <= 2^15 / (sc_in min se_v min se_p),  |dP| <= 64 max|dO| (max se_v qmax_v + max|v_aft|),
//   |dS se_k se_x| <= 2 alpha max|dP| max se_k max se_x;   sc = 2^k with bound * sc in [2^14, 2^15).
// The bound is loose (a few binades): with 5 exponent bits that costs range at the bottom, where gradients do not matter.
__global__ void __launch_bounds__(256)
attn_bwd_scale_kernel(const float* __restrict__ sc_in, const float* __restrict__ se_v, const float* __restrict__ v_aft, int C,
                      const float* __restrict__ se_p, const float* __restrict__ se_x, int N, const float* __restrict__ se_k, int NH,
                      float qmax_v, float alpha, float* __restrict__ out2) {
    __shared__ float red[6][8];
    float mn_v = INFINITY, mx_v = 0.f, mx_a = 0.f, mn_p = INFINITY, mx_x = 0.f, mx_k = 0.f;
    for (int i = threadIdx.x; i < C; i += 256) { const float a = fabsf(se_v[i]); mn_v = fminf(mn_v, a); mx_v = fmaxf(mx_v, a); mx_a = fmaxf(mx_a, fabsf(v_aft[i])); }
    for (int i = threadIdx.x; i < N; i += 256) { mn_p = fminf(mn_p, fabsf(se_p[i])); mx_x = fmaxf(mx_x, fabsf(se_x[i])); }
    for (int i = threadIdx.x; i < NH; i += 256) mx_k = fmaxf(mx_k, fabsf(se_k[i]));
    float v[6] = {-mn_v, mx_v, mx_a, -mn_p, mx_x, mx_k};          // minima as maxima of the negated values
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 6; ++k) { float a = red[k][0]; for (int w = 1; w < 8; ++w) a = fmaxf(a, red[k][w]); v[k] = a; }
        const float max_dO = 32768.f / (sc_in[0] * fmaxf(-v[0], 1e-30f) * fmaxf(-v[3], 1e-30f));
        const float max_dP = 64.f * max_dO * (v[1] * qmax_v + v[2]);
        float bound = 2.f * alpha * max_dP * v[5] * v[4];
        if (!(bound > 0.f) || !isfinite(bound)) bound = 1.f;
        int e;
        frexpf(bound, &e);                            // bound = f * 2^e, f in [0.5, 1)  ->  bound * 2^(15 - e) in [2^14, 2^15)
        const float sc = ldexpf(1.f, 15 - e);
        out2[0] = sc;
        out2[1] = 1.f / sc;
    }
}

// d_s[n] = g_s * sum_z part[z][n] in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
attn_ds_reduce_kernel(const float* __restrict__ part, int nz, int N, float g_s, float* __restrict__ d_s) {
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

}  // namespace attn
}  // namespace ofq

using namespace ofq;

static int make_map_u8(CUtensorMap* tm, const void* ptr, const cuuint64_t (&dims)[5], const cuuint64_t (&strides)[4],
                       cuuint32_t box_rows) {
    cuuint32_t box[5] = {128, box_rows, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return ofq_encode_tensor_map(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(ptr), dims, strides, box, estr);
}

extern "C" int ofq_qkr_attn_fwd(const int8_t* qx, const int8_t* qk, const int8_t* qvT, long long ldv, int B, int N, int H, int C,
                                const float* se_x, const float* se_k, const float* ctS, float scale, const float* se_p,
                                int qhi, const float* se_v, const float* v_aft, int8_t* qp, long long ldq, float* out,
                                float* P, long long ldS, void* qp16, int fmt16, float* rowsum, float* rowstat, void* stream) {
    OFQ_REQUIRE(qx && qk && qvT && se_x && se_k && ctS && se_p && se_v && v_aft && qp && out, "ofq_qkr_attn_fwd: null argument");
    OFQ_REQUIRE(B > 0 && H > 0 && N > 0 && N <= attn::BN && C == H * attn::HD && C % 16 == 0 && C <= 128 * attn::MAXKB,
                "ofq_qkr_attn_fwd: needs head dim 64, at most 208 tokens and C <= 384 (got N=%d H=%d C=%d)", N, H, C);
    OFQ_REQUIRE(ldq >= attn::BN && ldq % 16 == 0 && ldv % 16 == 0 && ldv >= N && (!P || (ldS % 4 == 0 && ldS >= N)),
                "ofq_qkr_attn_fwd: code pitch must be a multiple of 16 and >= 208, probability pitch a multiple of 4");
    OFQ_REQUIRE(((uintptr_t)qx | (uintptr_t)qk | (uintptr_t)qvT | (uintptr_t)qp | (uintptr_t)out | (uintptr_t)P | (uintptr_t)qp16) % 16 == 0,
                "ofq_qkr_attn_fwd: tensors must be 16-byte aligned");
    OFQ_REQUIRE(qhi > 0 && qhi <= 127, "ofq_qkr_attn_fwd: bad level count");
    OFQ_REQUIRE(!qp16 || fmt16 == OFQ_FMT_BF16 || fmt16 == OFQ_FMT_F16, "ofq_qkr_attn_fwd: bad 16-bit format");
    OFQ_CHECK_ARCH();
    CUtensorMap tmQ, tmK, tmV;
    {   // x_hat codes [B, N, C]
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B, 1, 1};
        const cuuint64_t st[4] = {(cuuint64_t)C, (cuuint64_t)N * C, (cuuint64_t)N * C, (cuuint64_t)N * C};
        int rc = make_map_u8(&tmQ, qx, dims, st, attn::BM);
        if (rc) return rc;
    }
    {   // k_hat codes [B, N, H, C]: rows = keys d (stride H*C), then head, then batch
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B, 1};
        const cuuint64_t st[4] = {(cuuint64_t)H * C, (cuuint64_t)C, (cuuint64_t)N * H * C, (cuuint64_t)N * H * C};
        int rc = make_map_u8(&tmK, qk, dims, st, attn::BN);
        if (rc) return rc;
    }
    {   // v_hat codes transposed [B, C, ldv] (keys contiguous, zero padded)
        const cuuint64_t dims[5] = {(cuuint64_t)ldv, (cuuint64_t)C, (cuuint64_t)B, 1, 1};
        const cuuint64_t st[4] = {(cuuint64_t)ldv, (cuuint64_t)C * ldv, (cuuint64_t)C * ldv, (cuuint64_t)C * ldv};
        int rc = make_map_u8(&tmV, qvT, dims, st, attn::HD);
        if (rc) return rc;
    }
    CUtensorMap tmP;
    {   // probability codes [B*H, N, ldq]: written by TMA from the shared-memory P operand
        const cuuint64_t dims[5] = {(cuuint64_t)ldq, (cuuint64_t)N, (cuuint64_t)B * H, 1, 1};
        const cuuint64_t st[4] = {(cuuint64_t)ldq, (cuuint64_t)N * ldq, (cuuint64_t)N * ldq, (cuuint64_t)N * ldq};
        int rc = make_map_u8(&tmP, qp, dims, st, attn::BM);
        if (rc) return rc;
    }
    attn::Params p;
    p.B = B; p.N = N; p.H = H; p.C = C; p.kblocks = (C + 127) / 128; p.units = B * H;
    p.se_x = se_x; p.se_k = se_k; p.ctS = ctS; p.scale = scale; p.se_p = se_p; p.qhi = (float)qhi;
    p.se_v = se_v; p.v_aft = v_aft; p.qp = qp; p.ldq = ldq; p.out = out; p.P = P; p.ldS = P ? ldS : 0;
    static const int dbg = [] { const char* e = getenv("OFQ_ATTN_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = dbg;
    p.qp16 = (uint16_t*)qp16; p.f16 = fmt16 == OFQ_FMT_F16; p.rowsum = rowsum; p.rowstat = reinterpret_cast<float2*>(rowstat);
    static bool configured = false;
    if (!configured) {
        OFQ_CUDA(cudaFuncSetAttribute(attn::qkr_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn::DYN_BYTES));
        configured = true;
    }
    const int grid = p.units < ofq_num_sms() ? p.units : ofq_num_sms();
    // 16 softmax warps by default; OFQ_ATTN_WARPS=8 keeps the first (8-warp) kernel for A/B measurements
    static const int warps = [] { const char* e = getenv("OFQ_ATTN_WARPS"); return e ? atoi(e) : 16; }();
    if (warps == 8) {
        attn::qkr_attn_fwd_kernel<<<grid, attn::NTHREADS, attn::DYN_BYTES, (cudaStream_t)stream>>>(tmQ, tmK, tmV, tmP, p);
    } else {
        static bool configured16 = false;
        if (!configured16) {
            OFQ_CUDA(cudaFuncSetAttribute(attn::qkr_attn_fwd16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn::DYN_BYTES16));
            configured16 = true;
        }
        attn::qkr_attn_fwd16_kernel<<<grid, attn::NTHREADS16, attn::DYN_BYTES16, (cudaStream_t)stream>>>(tmQ, tmK, tmV, tmP, p);
    }
    OFQ_CUDA(cudaGetLastError());
    return 0;
}


static int make_map_16(CUtensorMap* tm, const void* ptr, const cuuint64_t (&dims)[5], const cuuint64_t (&strides)[4],
                       cuuint32_t box_rows, bool f16) {
    cuuint32_t box[5] = {64, box_rows, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return ofq_encode_tensor_map(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                                 const_cast<void*>(ptr), dims, strides, box, estr);
}

extern "C" int ofq_qkr_attn_bwd(const int8_t* qx, const int8_t* qk, const void* a16, const void* qv16, int fmt16, int B, int N,
                                int H, int C, const float* se_x, const float* se_k, const float* ctS, float scale,
                                const float* se_p, const float* inv_se_p, int qhi, const float* rowstat, const float* rowdot,
                                const float* sc_in, const float* se_v, const float* v_aft, int qmax_v, float g_s, void* dS16,
                                long long ldo, float* colsum, float* ds_part, float* d_s, float* sc_out, void* stream) {
    OFQ_REQUIRE(qx && qk && a16 && qv16 && se_x && se_k && ctS && se_p && inv_se_p && rowstat && rowdot && sc_in && se_v && v_aft &&
                dS16 && colsum && ds_part && d_s && sc_out, "ofq_qkr_attn_bwd: null argument");
    OFQ_REQUIRE(B > 0 && H > 0 && N > 0 && N <= attn::BN && C == H * attn::HD && C <= 128 * attn::MAXKB,
                "ofq_qkr_attn_bwd: needs head dim 64, at most 208 tokens and C <= 384 (got N=%d H=%d C=%d)", N, H, C);
    OFQ_REQUIRE(fmt16 == OFQ_FMT_BF16 || fmt16 == OFQ_FMT_F16, "ofq_qkr_attn_bwd: bad 16-bit format");
    OFQ_REQUIRE(ldo % 8 == 0 && ldo >= N, "ofq_qkr_attn_bwd: output pitch must be a multiple of 8 and >= N");
    OFQ_REQUIRE(((uintptr_t)qx | (uintptr_t)qk | (uintptr_t)a16 | (uintptr_t)qv16 | (uintptr_t)dS16) % 16 == 0 && (uintptr_t)rowstat % 8 == 0,
                "ofq_qkr_attn_bwd: tensors must be 16-byte aligned");
    OFQ_CHECK_ARCH();
    const bool f16 = fmt16 == OFQ_FMT_F16;
    CUtensorMap tmQ, tmK, tmA, tmV;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B, 1, 1};
        const cuuint64_t st[4] = {(cuuint64_t)C, (cuuint64_t)N * C, (cuuint64_t)N * C, (cuuint64_t)N * C};
        int rc = make_map_u8(&tmQ, qx, dims, st, attn::BM);
        if (rc) return rc;
    }
    {
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B, 1};
        const cuuint64_t st[4] = {(cuuint64_t)H * C, (cuuint64_t)C, (cuuint64_t)N * H * C, (cuuint64_t)N * H * C};
        int rc = make_map_u8(&tmK, qk, dims, st, attn::BN);
        if (rc) return rc;
    }
    {   // 16-bit tensors [B, N, C]: one 64-channel (128-byte) slice of a head per box row
        const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B, 1, 1};
        const cuuint64_t st[4] = {(cuuint64_t)C * 2, (cuuint64_t)N * C * 2, (cuuint64_t)N * C * 2, (cuuint64_t)N * C * 2};
        int rc = make_map_16(&tmA, a16, dims, st, attn::BM, f16);
        if (rc) return rc;
        rc = make_map_16(&tmV, qv16, dims, st, attn::BN, f16);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    attn::attn_bwd_scale_kernel<<<1, 256, 0, st>>>(sc_in, se_v, v_aft, C, se_p, se_x, N, se_k, N * H, (float)qmax_v, scale, sc_out);
    attn::BwdParams p;
    p.B = B; p.N = N; p.H = H; p.C = C; p.kblocks = (C + 127) / 128; p.units = B * H;
    p.se_x = se_x; p.se_k = se_k; p.ctS = ctS; p.scale = scale; p.se_p = se_p; p.inv_se_p = inv_se_p; p.qhi = (float)qhi;
    p.rowstat = reinterpret_cast<const float2*>(rowstat); p.rowdot = rowdot; p.sc_in = sc_in; p.sc_out = sc_out;
    p.dS16 = (uint16_t*)dS16; p.ldo = ldo; p.f16 = f16; p.colsum = colsum; p.ds_part = ds_part;
    static bool configured = false;
    if (!configured) {
        OFQ_CUDA(cudaFuncSetAttribute(attn::qkr_attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn::B_DYN_BYTES));
        OFQ_CUDA(cudaFuncSetAttribute(attn::qkr_attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn::B_DYN_BYTES));
        configured = true;
    }
    const int grid = p.units < ofq_num_sms() ? p.units : ofq_num_sms();
    if (f16) attn::qkr_attn_bwd_kernel<true><<<grid, attn::NTHREADS16, attn::B_DYN_BYTES, st>>>(tmQ, tmK, tmA, tmV, p);
    else attn::qkr_attn_bwd_kernel<false><<<grid, attn::NTHREADS16, attn::B_DYN_BYTES, st>>>(tmQ, tmK, tmA, tmV, p);
    attn::attn_ds_reduce_kernel<<<(N + 31) / 32, 256, 0, st>>>(ds_part, B * H, N, g_s, d_s);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}
