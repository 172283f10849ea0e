// Tensor-core GEMM engine of the OFQ hot path (sm_100a only).
//
//   D[z][m][n] (=|+=)  acc[z][m][n] * rs[m] * cs[n]  +  rt[m] * ct[n]
//   acc[z][m][n] = sum_{k2,k} A[z][k2][m][k] * B[z][k2][n][k]        (both operands K-major)
//
// kind::i8  : int8 quantization codes x int8 codes -> exact int32 accumulators in TMEM (forward path:
//             QLinear / qkx / attention-score / P.V GEMMs of reference qlinear.py:69, attention.py:180,200,210,219).
// kind::f16 : bf16 x bf16 -> fp32 accumulators (backward dX / dW / attention gradients).
//
// One CTA = one 128 x BN output tile. Warp 0 streams 128-byte-swizzled operand tiles with TMA into a
// multi-stage smem ring, warp 1 (one elected lane) issues tcgen05.mma into a TMEM accumulator, warps 2..5
// read the accumulator back with tcgen05.ld and apply the rank-1 scale/offset epilogue straight to HBM.
// Quantizer scales never touch the operands: per-row (token) scales, per-column (StatsQ channel) scales
// and the affine "move_aft" shift all live in the epilogue vectors, so the MMA itself is exact.
#include <cstdio>
#include <cstdlib>
#include "ofq_b200.h"
#include "ptx.cuh"
#include "host_util.h"

namespace ofq {

constexpr int BM = 128;          // UMMA M
constexpr int KBYTES = 128;      // one 128B swizzle atom of K per stage
constexpr int EPI_WARPS = 8;      // two warps per TMEM lane quarter, interleaved over the 32-column chunks
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32; // warp0 TMA, warp1 MMA, warps 2..9 epilogue

// Division / remainder of a non-negative int (< 2^31) by a run-time constant as multiply-high + shift: the tile decode and
// the wrapped epilogue-vector indices sit on the serial path of every tile, where a hardware-emulated integer division
// costs ~100 dependent cycles each.  q = (umulhi(n, mul) + n) >> shr with mul = floor(2^32 (2^shr - d) / d) + 1.
struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ int div(int n) const { return (int)((__umulhi((uint32_t)n, mul) + (uint32_t)n) >> shr); }
    __device__ __forceinline__ int mod(int n) const { return n - div(n) * (int)d; }
};
static FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (uint32_t)(d < 1 ? 1 : d);
    uint32_t s = 0;
    while (s < 31 && (1u << s) < f.d) ++s;
    f.shr = s;
    f.mul = (uint32_t)((((unsigned long long)1 << 32) * (((unsigned long long)1 << s) - f.d)) / f.d + 1);
    return f;
}

// 32-column chunks of a BN-wide tile owned by one epilogue warp (the two warps of a lane quarter take even / odd chunks)
constexpr int epi_nch(int bn) { return (bn / 32 + 1) / 2; }

struct VecRef {
    const float* p;   // nullptr -> 1.0
    int period;       // rs / rt / cs: index = i % period (ct is never wrapped)
    long long bs1, bs2;
    FastDiv fd;       // of `period`
};

struct GemmParams {
    int M, N;
    int kblocks, k2, splits;
    int nb1, nb2;
    int a_b1, a_b2, b_b1, b_b2, c_b1, c_b2;
    int a_k2, b_k2;          // 0: operand shared by all outer-K slices
    int a_k2mod, b_k2mod;    // slice index = k2 % mod
    int a_dual_delta;        // dual-A: second A tile of a stage comes from outer-K slice (k2 + delta)
    VecRef rs, cs, rt, ct;
    int has_rank1;
    int atomic;
    int ab_fmt;              // 16-bit kinds: UMMA a/b format field (0 = F16, 1 = BF16)
    int a_mn, b_mn;          // 16-bit kinds: operand is MN-major (rows contiguous, K strided) instead of K-major
    unsigned int* amax;      // optional: max |output| over the whole problem (bits of a non-negative float, atomicMax)
    FastDiv fd_ntiles, fd_mtiles, fd_nbatch, fd_nb1, fd_splits, fd_kblocks, fd_ak2mod, fd_bk2mod;   // set by the launchers
    int debug_nostore;       // measurement only (OFQ_GEMM_NOSTORE bits): 1 = epilogue without stores, 2 = no operand loads, 4 = no epilogue work (single-CTA kernel)
};

// NA: A tiles per pipeline stage. NA = 2 ("dual-A") loads the bf16 hi and lo planes of a gradient operand together with
// ONE copy of the shared code operand B, instead of streaming B twice: a third less L2 -> SM traffic per MAC.
// OUT_BUFS: TMA-store staging buffers per epilogue warp (stores in flight).
template <int BN, int STAGES, int NA, int OUT_BUFS>
struct SmemLayout {
    static constexpr uint32_t A_BYTES = BM * KBYTES;
    static constexpr uint32_t B_BYTES = BN * KBYTES;
    static constexpr uint32_t STAGE_BYTES = NA * A_BYTES + B_BYTES;
    static constexpr uint32_t OUT_OFF = STAGES * STAGE_BYTES;          // 4 warps x OUT_BUFS x (32 x 128 B), swizzled
    static constexpr uint32_t OUT_BYTES = EPI_WARPS * OUT_BUFS * 4096;
    static constexpr uint32_t VEC_OFF = OUT_OFF + OUT_BYTES;            // per epilogue warp: {cs[32], ct[32]} per owned chunk
    static constexpr uint32_t BAR_OFF = VEC_OFF + EPI_WARPS * epi_nch(BN) * 64 * 4;
    static constexpr uint32_t TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr size_t DYN_BYTES = TOTAL + 1024;                   // slack for 1024 B alignment
};

struct TileCoord {
    int m0, n0, b1, b2, split, it_begin, nit;
};

// tile index -> coordinates; consecutive indices share the A (row) block so that it is re-read from L2
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int t, int BN_, int mtiles, int ntiles) {
    TileCoord c;
    int q = p.fd_ntiles.div(t);
    const int nb = t - q * ntiles; t = q;
    q = p.fd_mtiles.div(t);
    const int mb = t - q * mtiles; t = q;
    c.split = p.fd_nbatch.div(t);
    const int z = t - c.split * (p.nb1 * p.nb2);
    c.b2 = p.fd_nb1.div(z); c.b1 = z - c.b2 * p.nb1;
    c.m0 = mb * BM; c.n0 = nb * BN_;
    const int total_it = p.k2 * p.kblocks;            // total_it * (splits + 1) < 2^31 (checked on the host)
    c.it_begin = p.fd_splits.div(total_it * c.split);
    c.nit = p.fd_splits.div(total_it * (c.split + 1)) - c.it_begin;
    return c;
}

// OFQ_GEMM_TRACE (measurement builds only, OFQ_NVCC_FLAGS=-DOFQ_GEMM_TRACE): per-phase clock64 totals of the first epilogue
// warp and the MMA thread of CTA 0, printed at kernel exit.
#ifdef OFQ_GEMM_TRACE
#define TR_DECL(n) long long n = 0
#define TR_T0(t) const long long t = clock64()
#define TR_ADD(n, t) n += clock64() - t
#else
#define TR_DECL(n)
#define TR_T0(t)
#define TR_ADD(n, t)
#endif
struct EpiTrace { long long ldwait, rdwait, math, fence; };

// ------------------------------------------------------------------------------------------ epilogue (shared)
// One 32 x 32 chunk of the accumulator: o = acc * rs[m] * cs[n] + rt[m] * ct[n] -> 128B-swizzled staging row `lane`.
// R1 / TRACK are warp-uniform facts lifted to template parameters: without the rank-1 term the ct vector is never read,
// without an amax request no maxima are formed; the register buffer is a compile-time array, so the eight float4 groups
// carry no branches and the scheduler can interleave them (the epilogue is latency-bound: 2 warps per scheduler).
template <int KIND, bool R1, bool TRACK>
__device__ __forceinline__ void epi_math(const uint32_t (&r)[32], const float rsv, const float rtv,
                                         const float* __restrict__ cs_c, const float* __restrict__ ct_c,
                                         float4* __restrict__ rowp, const int lane, float& omax) {
    // two columns per instruction (FMUL2 / FFMA2): same roundings as fmaf(acc * rs, cs, rt * ct) per element.
    // Shared-memory traffic goes through volatile asm without memory clobbers, in two batches of four 16-byte groups: all
    // vector loads of a batch are issued before its first staging store. As plain C++ accesses every store would pin the
    // following loads behind it (neither nvvm nor ptxas can prove that the staging buffer and the vectors do not alias),
    // which serialised the eight groups on the shared-memory latency. The fence.proxy.async after the chunk is a volatile
    // asm with a memory clobber and stays behind these stores.
    const float2 rs2 = make_float2(rsv, rsv), rt2 = make_float2(rtv, rtv);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float4 cs4[4], ct4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            cs4[g] = ld_shared_v4_nc(cs_c + 16 * b + 4 * g);
            if (R1) ct4[g] = ld_shared_v4_nc(ct_c + 16 * b + 4 * g);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int j4 = 4 * b + g;
            float2 o[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t ra = r[4 * j4 + 2 * h], rb = r[4 * j4 + 2 * h + 1];
                const float2 acc = KIND == 0 ? make_float2(static_cast<float>(static_cast<int32_t>(ra)), static_cast<float>(static_cast<int32_t>(rb)))
                                             : make_float2(__uint_as_float(ra), __uint_as_float(rb));
                const float2 cs2 = h == 0 ? make_float2(cs4[g].x, cs4[g].y) : make_float2(cs4[g].z, cs4[g].w);
                const float2 t = __fmul2_rn(acc, rs2);
                if (R1) {
                    const float2 ct2 = h == 0 ? make_float2(ct4[g].x, ct4[g].y) : make_float2(ct4[g].z, ct4[g].w);
                    o[h] = __ffma2_rn(t, cs2, __fmul2_rn(rt2, ct2));
                } else {
                    o[h] = __fmul2_rn(t, cs2);
                }
                if (TRACK) omax = fmaxf(omax, fmaxf(fabsf(o[h].x), fabsf(o[h].y)));   // rows / columns outside the matrix carry exact zeros
            }
            st_shared_v4_nc(rowp + (j4 ^ (lane & 7)), o[0].x, o[0].y, o[1].x, o[1].y);
        }
    }
}

// The epilogue vectors of one tile as per-lane registers: lane l holds cs / ct of column l of every chunk its warp owns,
// plus rs / rt of its row. Loaded one tile ahead (the global-load latency hides behind the chunk work of the current tile)
// and published to a warp-private staging area: no CTA-wide barrier, no coupling between the epilogue warps.
template <int NCH>
struct EpiVecs {
    float cs[NCH], ct[NCH];
    float rsv, rtv;
    bool rank1;
};
template <int NCH>
__device__ __forceinline__ void epi_vec_load(const GemmParams& p, const TileCoord& c, const int m, const int half,
                                             const int lane, EpiVecs<NCH>& v) {
    v.rank1 = p.has_rank1 && c.split == 0;
    const long long cs_off = (long long)c.b1 * p.cs.bs1 + (long long)c.b2 * p.cs.bs2;
    const long long ct_off = (long long)c.b1 * p.ct.bs1 + (long long)c.b2 * p.ct.bs2;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int n = c.n0 + (half + 2 * k) * 32 + lane;
        const bool ok = n < p.N;
        v.cs[k] = ok ? (p.cs.p ? __ldg(p.cs.p + cs_off + p.cs.fd.mod(n)) : 1.0f) : 0.f;
        v.ct[k] = (ok && v.rank1) ? (p.ct.p ? __ldg(p.ct.p + ct_off + n) : 1.0f) : 0.f;
    }
    const bool row_ok = m < p.M;
    const long long rs_off = (long long)c.b1 * p.rs.bs1 + (long long)c.b2 * p.rs.bs2;
    const long long rt_off = (long long)c.b1 * p.rt.bs1 + (long long)c.b2 * p.rt.bs2;
    v.rsv = row_ok ? (p.rs.p ? __ldg(p.rs.p + rs_off + p.rs.fd.mod(m)) : 1.0f) : 0.f;
    v.rtv = (row_ok && v.rank1) ? (p.rt.p ? __ldg(p.rt.p + rt_off + p.rt.fd.mod(m)) : 1.0f) : 0.f;
}
template <int NCH>
__device__ __forceinline__ void epi_vec_publish(const EpiVecs<NCH>& v, float* myvec, const int lane) {
    __syncwarp();                      // every lane is done reading the previous tile's vectors
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        myvec[k * 64 + lane] = v.cs[k];
        myvec[k * 64 + 32 + lane] = v.ct[k];
    }
    __syncwarp();
}

// All chunks a warp owns of one output tile: TMEM loads one chunk ahead into two compile-time register buffers, scale /
// offset math, swizzled staging, TMA store (or reduce-add). `have` = the tile has an accumulator at all (nit > 0).
template <int KIND, int OUT_BUFS>
__device__ __forceinline__ void epi_tile(const CUtensorMap* tmC, const GemmParams& p, const TileCoord& c, const int m_row0,
                                         const uint32_t tmem_acc, const bool have, const int nvalid, const int half,
                                         const bool rank1, const float rsv, const float rtv, const float* myvec,
                                         uint8_t* stage_base, uint32_t& chunk, const int lane,
                                         float& omax, EpiTrace* tr = nullptr) {
    uint32_t r0[32], r1[32];
    const bool track = p.amax != nullptr;
    auto fetch = [&](int cc, uint32_t (&rr)[32]) {
        if (have) {
            tmem_ld_32x32(tmem_acc + cc * 32, rr);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = 0u;
        }
    };
    auto emit = [&](const uint32_t (&rr)[32], int cc) {
        const int c0 = cc * 32;
        const float* cs_c = myvec + ((cc - half) >> 1) * 64;       // {cs[32], ct[32]} of this owned chunk
        const float* ct_c = cs_c + 32;
        uint8_t* buf = stage_base + (chunk % OUT_BUFS) * 4096;
        TR_T0(t_a);
        if (chunk >= OUT_BUFS) {  // the staging buffer used OUT_BUFS chunks ago must have been read by TMA
            if (lane == 0) tma_store_wait_read<OUT_BUFS - 1>();
            __syncwarp();
        }
        ++chunk;
#ifdef OFQ_GEMM_TRACE
        if (tr) TR_ADD(tr->rdwait, t_a);
#endif
        TR_T0(t_b);
        // row `lane` of the 32x32 fp32 box, 128B-swizzled: 16-byte chunk j lands at j ^ (lane % 8)
        float4* rowp = reinterpret_cast<float4*>(buf + lane * 128);
        if (rank1) {
            if (track) epi_math<KIND, true, true>(rr, rsv, rtv, cs_c, ct_c, rowp, lane, omax);
            else       epi_math<KIND, true, false>(rr, rsv, rtv, cs_c, ct_c, rowp, lane, omax);
        } else {
            if (track) epi_math<KIND, false, true>(rr, rsv, rtv, cs_c, ct_c, rowp, lane, omax);
            else       epi_math<KIND, false, false>(rr, rsv, rtv, cs_c, ct_c, rowp, lane, omax);
        }
#ifdef OFQ_GEMM_TRACE
        if (tr) TR_ADD(tr->math, t_b);
#endif
        TR_T0(t_c);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.debug_nostore & 1)) {
            if (p.atomic)
                tma_reduce_add_5d(tmC, buf, c.n0 + c0, m_row0, 0, c.b1 * p.c_b1, c.b2 * p.c_b2);
            else
                tma_store_5d(tmC, buf, c.n0 + c0, m_row0, 0, c.b1 * p.c_b1, c.b2 * p.c_b2);
            tma_store_commit();
        }
#ifdef OFQ_GEMM_TRACE
        if (tr) TR_ADD(tr->fence, t_c);
#endif
    };
    int ci = half;
    if (ci < nvalid) fetch(ci, r0);
#pragma unroll 1
    for (; ci < nvalid; ci += 4) {
        TR_T0(t_l0);
        if (have) { tmem_ld_wait(); tmem_ld_pin(r0); }        // chunk ci is in r0
#ifdef OFQ_GEMM_TRACE
        if (tr) TR_ADD(tr->ldwait, t_l0);
#endif
        if (ci + 2 < nvalid) fetch(ci + 2, r1);               // prefetch the next owned chunk
        emit(r0, ci);
        if (ci + 2 < nvalid) {
            TR_T0(t_l1);
            if (have) { tmem_ld_wait(); tmem_ld_pin(r1); }
#ifdef OFQ_GEMM_TRACE
            if (tr) TR_ADD(tr->ldwait, t_l1);
#endif
            if (ci + 4 < nvalid) fetch(ci + 4, r0);
            emit(r1, ci + 2);
        }
    }
}

// Persistent, warp-specialised: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The smem ring runs
// across tile boundaries and the TMEM accumulator is double buffered, so the epilogue of tile i overlaps the TMA
// loads and MMAs of tile i+1.
template <int KIND, int BN, int STAGES, int NA, int OUT_BUFS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p, const int num_tiles,
               const int mtiles, const int ntiles) {
    using L = SmemLayout<BN, STAGES, NA, OUT_BUFS>;
    constexpr uint32_t A_BYTES = L::A_BYTES;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES;
    constexpr uint32_t ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));   // per accumulator stage
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    constexpr uint32_t UMMA_K_BYTES = 32;  // 32 int8 or 16 bf16 per MMA
    // S32 accumulate / signed int8, or F32 accumulate / {fp16, bf16} (format chosen at run time)
    const uint32_t IDESC = KIND == 0 ? umma_idesc(2u, 1u, BM, BN)
                                     : (umma_idesc(1u, (uint32_t)p.ab_fmt, BM, BN) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16));

    // 1024-byte alignment is required by the 128B swizzle; the attribute keeps the pointer in the shared address
    // space (an integer round-up would degrade every access to generic LD/ST)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;     // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;          // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* vec_s = reinterpret_cast<float*>(smem + L::VEC_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], EPI_WARPS); // one arrival per epilogue warp
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

    if (warp == 0) {
        if (elect_one()) {
            const int kelem = KIND == 0 ? KBYTES : KBYTES / 2;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int g = c.it_begin + i;
                    const int k2i = p.fd_kblocks.div(g), kb = g - k2i * p.kblocks;
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    uint8_t* sb = sa + NA * A_BYTES;
                    if (p.debug_nostore & 2) { mbar_arrive(&full_bar[s]); continue; }   // probe: no operand loads
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    if (KIND != 0 && p.a_mn) {   // MN-major: 64-row x 64-k boxes, one 8 KB swizzle-atom column each
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_5d(sa + j * 8192, &tmA, &full_bar[s], c.m0 + 64 * j, kb * kelem, p.fd_ak2mod.mod(k2i) * p.a_k2, c.b1 * p.a_b1, c.b2 * p.a_b2);
                    } else {
                        tma_load_5d(sa, &tmA, &full_bar[s], kb * kelem, c.m0, p.fd_ak2mod.mod(k2i) * p.a_k2, c.b1 * p.a_b1, c.b2 * p.a_b2);
                        if (NA == 2)
                            tma_load_5d(sa + A_BYTES, &tmA, &full_bar[s], kb * kelem, c.m0, k2i + p.a_dual_delta, c.b1 * p.a_b1, c.b2 * p.a_b2);
                    }
                    if (KIND != 0 && p.b_mn) {
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_5d(sb + j * 8192, &tmB, &full_bar[s], c.n0 + 64 * j, kb * kelem, p.fd_bk2mod.mod(k2i) * p.b_k2, c.b1 * p.b_b1, c.b2 * p.b_b2);
                    } else {
                        tma_load_5d(sb, &tmB, &full_bar[s], kb * kelem, c.n0, p.fd_bk2mod.mod(k2i) * p.b_k2, c.b1 * p.b_b1, c.b2 * p.b_b2);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            uint32_t it = 0, tc = 0;
            TR_DECL(tm_acc); TR_DECL(tm_full);
            TR_T0(tm_begin);
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
                TR_T0(t_e);
                mbar_wait(&acc_empty[as], aph ^ 1);       // epilogue has drained this accumulator stage
                TR_ADD(tm_acc, t_e);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    TR_T0(t_f);
                    mbar_wait(&full_bar[s], ph);
                    TR_ADD(tm_full, t_f);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    // K-major: one MMA consumes 32 bytes of the 128-byte K row. MN-major (16-bit): 16 k-rows of 128 bytes.
                    const bool bmn = KIND != 0 && p.b_mn, amn = KIND != 0 && p.a_mn;
                    const uint64_t bdesc = bmn ? umma_desc_mnmajor_sw128(sa + NA * A_BYTES) : umma_desc_kmajor_sw128(sa + NA * A_BYTES);
                    const uint64_t badv = bmn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
                    const uint64_t aadv = amn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
#pragma unroll
                    for (uint32_t na = 0; na < (uint32_t)NA; ++na) {
                        const uint64_t adesc = amn ? umma_desc_mnmajor_sw128(sa + na * A_BYTES) : umma_desc_kmajor_sw128(sa + na * A_BYTES);
#pragma unroll
                        for (uint32_t kk = 0; kk < KBYTES / UMMA_K_BYTES; ++kk) {
                            if (KIND == 0)
                                umma_i8(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk | na) != 0);
                            else
                                umma_f16(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk | na) != 0);
                        }
                    }
                    tc_commit(&empty_bar[s]);      // frees the smem stage when these MMAs retire
                }
                tc_commit(&acc_full[as]);          // accumulator of this tile complete
            }
#ifdef OFQ_GEMM_TRACE
            if (blockIdx.x == 0)
                printf("mma thread: tiles %u total %lld | wait acc_empty %lld wait full %lld\n", tc, clock64() - tm_begin, tm_acc, tm_full);
#endif
        }
    } else {
        // ---- epilogue: warps 2..9. TMEM lane quarter = warp % 4 (hardware restriction); the two warps of a quarter
        //      take the even / odd 32-column chunks. TMEM loads are software-pipelined one chunk ahead.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        uint8_t* stage_base = smem + L::OUT_OFF + (warp - 2) * OUT_BUFS * 4096;
        uint32_t tc = 0, chunk = 0;
        float omax = 0.f;                          // max |output| seen by this thread (p.amax)
        TR_DECL(tr_vec); TR_DECL(tr_wait); TR_DECL(tr_work);
        EpiTrace etr = {0, 0, 0, 0};
        TR_T0(tr_begin);
        constexpr int NCH = epi_nch(BN);
        float* myvec = vec_s + (warp - 2) * NCH * 64;
        EpiVecs<NCH> nv;                           // vectors of the NEXT tile, loaded one tile ahead
        TileCoord cn;
        if ((int)blockIdx.x < num_tiles) {
            cn = decode_tile(p, blockIdx.x, BN, mtiles, ntiles);
            epi_vec_load<NCH>(p, cn, cn.m0 + q * 32 + lane, half, lane, nv);
        }
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
            TR_T0(t_v);
            const TileCoord c = cn;
            const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
            const bool rank1 = nv.rank1;
            const float rsv = nv.rsv, rtv = nv.rtv;
            epi_vec_publish<NCH>(nv, myvec, lane);
            if (t + (int)gridDim.x < num_tiles) {
                cn = decode_tile(p, t + gridDim.x, BN, mtiles, ntiles);
                epi_vec_load<NCH>(p, cn, cn.m0 + q * 32 + lane, half, lane, nv);
            }

            TR_ADD(tr_vec, t_v);
            TR_T0(t_w);
            if (c.nit > 0) {
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
            }
            TR_ADD(tr_wait, t_w);
            TR_T0(t_k);
            const uint32_t tmem_acc = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            // number of chunks this warp owns inside the valid column range (warp-uniform)
            int nvalid = (min(BN, p.N - c.n0) + 31) / 32;
            if (c.m0 + q * 32 >= p.M) nvalid = 0;      // this warp's 32 rows lie entirely below the matrix (ragged last row block)
            if (p.debug_nostore & 4) nvalid = 0;                               // probe: no epilogue work at all
            epi_tile<KIND, OUT_BUFS>(&tmC, p, c, c.m0 + q * 32, tmem_acc, c.nit > 0, nvalid, half, rank1, rsv, rtv, myvec,
                                     stage_base, chunk, lane, omax, &etr);
            TR_ADD(tr_work, t_k);
            // all TMEM reads of this tile are complete: hand the accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (p.amax) {
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o2));
            if (lane == 0) atomicMax(p.amax, __float_as_uint(omax));
        }
#ifdef OFQ_GEMM_TRACE
        if (blockIdx.x == 0 && threadIdx.x == 64)
            printf("epi warp: tiles %u total %lld | vec %lld accwait %lld work %lld | ldwait %lld rdwait %lld math %lld fence+store %lld\n",
                   tc, clock64() - tr_begin, tr_vec, tr_wait, tr_work, etr.ldwait, etr.rdwait, etr.math, etr.fence);
#endif
        if (lane == 0) tma_store_wait_all<0>();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


__device__ __forceinline__ void tmem_ld_32x32_x16r(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_pin16(uint32_t (&r)[16]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                      "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) : : "memory");
}
// ------------------------------------------------------------------------------------------ 16-epilogue-warp variant
// Same mainloop; SIXTEEN epilogue warps (four per TMEM lane quarter, chunk j of a tile goes to warp j % 4 of its quarter), TMEM
// read-out in 16-column pieces, one staging buffer per warp: 576 threads at <= 112 registers. The epilogue of this engine is
// bound by latency chains (TMEM load -> math -> staging -> fence -> TMA store), not by issue slots: twice the warps per scheduler
// hide twice the latency (the same observation that took the fused attention kernels from 8 to 16 softmax warps).
#ifndef OFQ_GEMM_EPI16_DEFAULT
#define OFQ_GEMM_EPI16_DEFAULT 0
#endif
constexpr int EPI_WARPS16 = 16;
constexpr int NUM_THREADS16 = 64 + EPI_WARPS16 * 32;
constexpr int epi_nch16(int bn) { return (bn / 32 + 3) / 4; }

template <int BN, int STAGES>
struct SmemLayout16 {
    static constexpr uint32_t A_BYTES = BM * KBYTES;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + BN * KBYTES;
    static constexpr uint32_t OUT_OFF = STAGES * STAGE_BYTES;
    static constexpr uint32_t OUT_BYTES = EPI_WARPS16 * 4096;
    static constexpr uint32_t VEC_OFF = OUT_OFF + OUT_BYTES;
    static constexpr uint32_t BAR_OFF = VEC_OFF + EPI_WARPS16 * epi_nch16(BN) * 64 * 4;
    static constexpr uint32_t TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr size_t DYN_BYTES = TOTAL + 1024;
};

// sixteen columns (half `sub` of a chunk) of row `lane` into the swizzled staging row
template <int KIND, bool R1, bool TRACK>
__device__ __forceinline__ void epi_math16(const uint32_t (&r)[16], const float rsv, const float rtv, const float* __restrict__ cs_c,
                                           const float* __restrict__ ct_c, uint8_t* __restrict__ rowp, const int lane, const int sub,
                                           float& omax) {
    const float2 rs2 = make_float2(rsv, rsv), rt2 = make_float2(rtv, rtv);
    float4 cs4[4], ct4[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        cs4[g] = ld_shared_v4_nc(cs_c + 16 * sub + 4 * g);
        if (R1) ct4[g] = ld_shared_v4_nc(ct_c + 16 * sub + 4 * g);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float2 o[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t ra = r[4 * g + 2 * h], rb = r[4 * g + 2 * h + 1];
            const float2 acc = KIND == 0 ? make_float2(static_cast<float>(static_cast<int32_t>(ra)), static_cast<float>(static_cast<int32_t>(rb)))
                                         : make_float2(__uint_as_float(ra), __uint_as_float(rb));
            const float2 cs2 = h == 0 ? make_float2(cs4[g].x, cs4[g].y) : make_float2(cs4[g].z, cs4[g].w);
            const float2 t = __fmul2_rn(acc, rs2);
            if (R1) {
                const float2 ct2 = h == 0 ? make_float2(ct4[g].x, ct4[g].y) : make_float2(ct4[g].z, ct4[g].w);
                o[h] = __ffma2_rn(t, cs2, __fmul2_rn(rt2, ct2));
            } else {
                o[h] = __fmul2_rn(t, cs2);
            }
            if (TRACK) omax = fmaxf(omax, fmaxf(fabsf(o[h].x), fabsf(o[h].y)));
        }
        st_shared_v4_nc(rowp + (((sub * 4 + g) ^ (lane & 7)) << 4), o[0].x, o[0].y, o[1].x, o[1].y);
    }
}

template <int KIND, int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS16, 1)
gemm_tc16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p, const int num_tiles, const int mtiles, const int ntiles) {
    using L = SmemLayout16<BN, STAGES>;
    constexpr uint32_t A_BYTES = L::A_BYTES;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES;
    constexpr uint32_t ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    constexpr uint32_t UMMA_K_BYTES = 32;
    const uint32_t IDESC = KIND == 0 ? umma_idesc(2u, 1u, BM, BN)
                                     : (umma_idesc(1u, (uint32_t)p.ab_fmt, BM, BN) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16));
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* vec_s = reinterpret_cast<float*>(smem + L::VEC_OFF);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS16); }
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

    if (warp == 0) {
        if (elect_one()) {
            const int kelem = KIND == 0 ? KBYTES : KBYTES / 2;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int g = c.it_begin + i;
                    const int k2i = p.fd_kblocks.div(g), kb = g - k2i * p.kblocks;
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    if (KIND != 0 && p.a_mn) {
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_5d(sa + j * 8192, &tmA, &full_bar[s], c.m0 + 64 * j, kb * kelem, p.fd_ak2mod.mod(k2i) * p.a_k2, c.b1 * p.a_b1, c.b2 * p.a_b2);
                    } else {
                        tma_load_5d(sa, &tmA, &full_bar[s], kb * kelem, c.m0, p.fd_ak2mod.mod(k2i) * p.a_k2, c.b1 * p.a_b1, c.b2 * p.a_b2);
                    }
                    if (KIND != 0 && p.b_mn) {
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_5d(sb + j * 8192, &tmB, &full_bar[s], c.n0 + 64 * j, kb * kelem, p.fd_bk2mod.mod(k2i) * p.b_k2, c.b1 * p.b_b1, c.b2 * p.b_b2);
                    } else {
                        tma_load_5d(sb, &tmB, &full_bar[s], kb * kelem, c.n0, p.fd_bk2mod.mod(k2i) * p.b_k2, c.b1 * p.b_b1, c.b2 * p.b_b2);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            uint32_t it = 0, tc = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const bool bmn = KIND != 0 && p.b_mn, amn = KIND != 0 && p.a_mn;
                    const uint64_t bdesc = bmn ? umma_desc_mnmajor_sw128(sa + A_BYTES) : umma_desc_kmajor_sw128(sa + A_BYTES);
                    const uint64_t adesc = amn ? umma_desc_mnmajor_sw128(sa) : umma_desc_kmajor_sw128(sa);
                    const uint64_t badv = bmn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
                    const uint64_t aadv = amn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
#pragma unroll
                    for (uint32_t kk = 0; kk < KBYTES / UMMA_K_BYTES; ++kk) {
                        if (KIND == 0) umma_i8(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk) != 0);
                        else umma_f16(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk) != 0);
                    }
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&acc_full[as]);
            }
        }
    } else {
        const int q = warp & 3;
        const int jq = (warp - 2) >> 2;                      // 0..3: which of the quarter's four warps
        constexpr int NCH = epi_nch16(BN);
        uint8_t* buf = smem + L::OUT_OFF + (warp - 2) * 4096;
        float* myvec = vec_s + (warp - 2) * NCH * 64;
        const bool track = p.amax != nullptr;
        uint32_t tc = 0, nstore = 0;
        float omax = 0.f;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
            const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
            const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
            const int m_row0 = c.m0 + q * 32, m = m_row0 + lane;
            const bool rank1 = p.has_rank1 && c.split == 0;
            const bool row_ok = m < p.M;
            const float rsv = row_ok ? (p.rs.p ? __ldg(p.rs.p + (long long)c.b1 * p.rs.bs1 + (long long)c.b2 * p.rs.bs2 + p.rs.fd.mod(m)) : 1.0f) : 0.f;
            const float rtv = (row_ok && rank1) ? (p.rt.p ? __ldg(p.rt.p + (long long)c.b1 * p.rt.bs1 + (long long)c.b2 * p.rt.bs2 + p.rt.fd.mod(m)) : 1.0f) : 0.f;
            int nvalid = (min(BN, p.N - c.n0) + 31) / 32;
            if (m_row0 >= p.M) nvalid = 0;
            const int nown = nvalid > jq ? (nvalid - jq + 3) >> 2 : 0;
            __syncwarp();
            {
                const long long cs_off = (long long)c.b1 * p.cs.bs1 + (long long)c.b2 * p.cs.bs2;
                const long long ct_off = (long long)c.b1 * p.ct.bs1 + (long long)c.b2 * p.ct.bs2;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int n = c.n0 + (jq + 4 * k) * 32 + lane;
                    const bool ok = n < p.N;
                    myvec[k * 64 + lane] = ok ? (p.cs.p ? __ldg(p.cs.p + cs_off + p.cs.fd.mod(n)) : 1.0f) : 0.f;
                    myvec[k * 64 + 32 + lane] = (ok && rank1) ? (p.ct.p ? __ldg(p.ct.p + ct_off + n) : 1.0f) : 0.f;
                }
            }
            __syncwarp();
            const bool have = c.nit > 0;
            if (have) {
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
            }
            const uint32_t tmem_acc = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t ra[16], rb[16];
            if (nown > 0 && have) tmem_ld_32x32_x16r(tmem_acc + (uint32_t)(jq * 32), ra);
#pragma unroll 1
            for (int k = 0; k < nown; ++k) {
                const int cc = jq + 4 * k;
                const float* cs_c = myvec + k * 64;
                const float* ct_c = cs_c + 32;
                if (nstore > 0) {                            // the staging buffer must have been read by the previous store
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                }
                uint8_t* rowp = buf + lane * 128;
                if (have) {
                    tmem_ld_wait(); tmem_ld_pin16(ra);
                    tmem_ld_32x32_x16r(tmem_acc + (uint32_t)(cc * 32 + 16), rb);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { ra[j] = 0u; rb[j] = 0u; }
                }
                if (rank1) { if (track) epi_math16<KIND, true, true>(ra, rsv, rtv, cs_c, ct_c, rowp, lane, 0, omax); else epi_math16<KIND, true, false>(ra, rsv, rtv, cs_c, ct_c, rowp, lane, 0, omax); }
                else       { if (track) epi_math16<KIND, false, true>(ra, rsv, rtv, cs_c, ct_c, rowp, lane, 0, omax); else epi_math16<KIND, false, false>(ra, rsv, rtv, cs_c, ct_c, rowp, lane, 0, omax); }
                if (have) {
                    tmem_ld_wait(); tmem_ld_pin16(rb);
                    if (k + 1 < nown) tmem_ld_32x32_x16r(tmem_acc + (uint32_t)((cc + 4) * 32), ra);
                }
                if (rank1) { if (track) epi_math16<KIND, true, true>(rb, rsv, rtv, cs_c, ct_c, rowp, lane, 1, omax); else epi_math16<KIND, true, false>(rb, rsv, rtv, cs_c, ct_c, rowp, lane, 1, omax); }
                else       { if (track) epi_math16<KIND, false, true>(rb, rsv, rtv, cs_c, ct_c, rowp, lane, 1, omax); else epi_math16<KIND, false, false>(rb, rsv, rtv, cs_c, ct_c, rowp, lane, 1, omax); }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (p.atomic) tma_reduce_add_5d(&tmC, buf, c.n0 + cc * 32, m_row0, 0, c.b1 * p.c_b1, c.b2 * p.c_b2);
                    else          tma_store_5d(&tmC, buf, c.n0 + cc * 32, m_row0, 0, c.b1 * p.c_b1, c.b2 * p.c_b2);
                    tma_store_commit();
                }
                ++nstore;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (p.amax) {
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o2));
            if (lane == 0) atomicMax(p.amax, __float_as_uint(omax));
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ CTA-pair variant
// Same pipeline on a 2-CTA cluster: one tcgen05.mma.cta_group::2 of M = 256 per instruction. CTA r of the pair owns rows
// [m0 + 128 r, m0 + 128 r + 128) of the 256 x BN tile (its accumulator lives in its own TMEM) and stages its own A rows
// plus HALF of the B tile (rows n0 + r BN/2 ...): (128 + BN/2) operand rows per k-block and CTA instead of (128 + BN), a
// third less L2 -> shared-memory traffic, which is what bounds these GEMMs (profiles/: ~9.5 of the ~12 TB/s the L2 slices
// can deliver at 33 % tensor-pipe activity). Protocol (cutlass sm100 2-SM kernels): both CTAs issue TMA, all transaction
// bytes are credited to the LEADER's stage barrier; the leader's elected thread issues the MMAs and multicasts its
// commits to both CTAs' empty / acc_full barriers; the epilogue warps of both CTAs arrive on the leader's acc_empty.
template <int BN, int STAGES, int OUT_BUFS>
struct SmemLayout2 {
    static constexpr uint32_t A_BYTES = BM * KBYTES;
    static constexpr uint32_t B_BYTES = (BN / 2) * KBYTES;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t OUT_OFF = STAGES * STAGE_BYTES;
    static constexpr uint32_t OUT_BYTES = EPI_WARPS * OUT_BUFS * 4096;
    static constexpr uint32_t VEC_OFF = OUT_OFF + OUT_BYTES;
    static constexpr uint32_t BAR_OFF = VEC_OFF + EPI_WARPS * epi_nch(BN) * 64 * 4;
    static constexpr uint32_t TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr size_t DYN_BYTES = TOTAL + 1024;
};

template <int KIND, int BN, int STAGES, int OUT_BUFS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmParams p, const int num_tiles,
                    const int mtiles, const int ntiles) {
    using L = SmemLayout2<BN, STAGES, OUT_BUFS>;
    constexpr uint32_t A_BYTES = L::A_BYTES;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES;
    constexpr uint32_t ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    constexpr uint32_t UMMA_K_BYTES = 32;
    constexpr int BNH = BN / 2;                       // B rows staged by each CTA
    const uint32_t IDESC = KIND == 0 ? umma_idesc(2u, 1u, 2 * BM, BN)
                                     : (umma_idesc(1u, (uint32_t)p.ab_fmt, 2 * BM, BN) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16));

    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* vec_s = reinterpret_cast<float*>(smem + L::VEC_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);               // leader's: one arrive.expect_tx + the bytes of both CTAs
            mbar_init(&empty_bar[s], 1);              // one multicast commit per use
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 2 * EPI_WARPS);  // leader's: the epilogue warps of both CTAs
        }
        fence_mbar_init();
    }
    cluster_sync_all();                               // barrier inits visible to the peer before any remote arrive / TMA
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const int kelem = KIND == 0 ? KBYTES : KBYTES / 2;
            uint32_t it = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);     // c.m0 = base of the 256-row pair tile
                const int m0 = 2 * c.m0 + (int)rank * BM;
                const int n0 = c.n0 + (int)rank * BNH;
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int g = c.it_begin + i;
                    const int k2i = p.fd_kblocks.div(g), kb = g - k2i * p.kblocks;
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    if (KIND != 0 && p.a_mn) {
                        for (int j = 0; j < BM / 64; ++j)
                            tma_load_5d_pair(sa + j * 8192, &tmA, bar, m0 + 64 * j, kb * kelem, p.fd_ak2mod.mod(k2i) * p.a_k2, c.b1 * p.a_b1, c.b2 * p.a_b2);
                    } else {
                        tma_load_5d_pair(sa, &tmA, bar, kb * kelem, m0, p.fd_ak2mod.mod(k2i) * p.a_k2, c.b1 * p.a_b1, c.b2 * p.a_b2);
                    }
                    if (KIND != 0 && p.b_mn) {
                        for (int j = 0; j < BNH / 64; ++j)
                            tma_load_5d_pair(sb + j * 8192, &tmB, bar, n0 + 64 * j, kb * kelem, p.fd_bk2mod.mod(k2i) * p.b_k2, c.b1 * p.b_b1, c.b2 * p.b_b2);
                    } else {
                        tma_load_5d_pair(sb, &tmB, bar, kb * kelem, n0, p.fd_bk2mod.mod(k2i) * p.b_k2, c.b1 * p.b_b1, c.b2 * p.b_b2);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            uint32_t it = 0, tc = 0;
            for (int t = pair; t < num_tiles; t += npairs, ++tc) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
                mbar_wait(&acc_empty[as], aph ^ 1);       // both CTAs' epilogues have drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const bool bmn = KIND != 0 && p.b_mn, amn = KIND != 0 && p.a_mn;
                    const uint64_t bdesc = bmn ? umma_desc_mnmajor_sw128(sa + A_BYTES) : umma_desc_kmajor_sw128(sa + A_BYTES);
                    const uint64_t adesc = amn ? umma_desc_mnmajor_sw128(sa) : umma_desc_kmajor_sw128(sa);
                    const uint64_t badv = bmn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
                    const uint64_t aadv = amn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
#pragma unroll
                    for (uint32_t kk = 0; kk < KBYTES / UMMA_K_BYTES; ++kk) {
                        if (KIND == 0)
                            umma_i8_pair(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk) != 0);
                        else
                            umma_f16_pair(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk) != 0);
                    }
                    tc_commit_pair(&empty_bar[s]);      // frees the stage in both CTAs when these MMAs retire
                }
                tc_commit_pair(&acc_full[as]);          // accumulators of both CTAs complete
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        uint8_t* stage_base = smem + L::OUT_OFF + (warp - 2) * OUT_BUFS * 4096;
        uint32_t tc = 0, chunk = 0;
        float omax = 0.f;                          // max |output| seen by this thread (p.amax)
        constexpr int NCH = epi_nch(BN);
        float* myvec = vec_s + (warp - 2) * NCH * 64;
        EpiVecs<NCH> nv;                           // vectors of the NEXT tile, loaded one tile ahead
        TileCoord cn;
        if (pair < num_tiles) {
            cn = decode_tile(p, pair, BN, mtiles, ntiles);
            epi_vec_load<NCH>(p, cn, 2 * cn.m0 + (int)rank * BM + q * 32 + lane, half, lane, nv);
        }
        for (int t = pair; t < num_tiles; t += npairs, ++tc) {
            const TileCoord c = cn;
            const int m0 = 2 * c.m0 + (int)rank * BM;
            const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
            const bool rank1 = nv.rank1;
            const float rsv = nv.rsv, rtv = nv.rtv;
            epi_vec_publish<NCH>(nv, myvec, lane);
            if (t + npairs < num_tiles) {
                cn = decode_tile(p, t + npairs, BN, mtiles, ntiles);
                epi_vec_load<NCH>(p, cn, 2 * cn.m0 + (int)rank * BM + q * 32 + lane, half, lane, nv);
            }

            if (c.nit > 0) {
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
            }
            const uint32_t tmem_acc = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            int nvalid = (min(BN, p.N - c.n0) + 31) / 32;
            if (m0 + q * 32 >= p.M) nvalid = 0;        // this warp's 32 rows (or the CTA's whole half) lie below the matrix
            epi_tile<KIND, OUT_BUFS>(&tmC, p, c, m0 + q * 32, tmem_acc, c.nit > 0, nvalid, half, rank1, rsv, rtv, myvec,
                                     stage_base, chunk, lane, omax);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[as]), 0));
        }
        if (p.amax) {
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o2));
            if (lane == 0) atomicMax(p.amax, __float_as_uint(omax));
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    // no CTA of the pair may leave (or free its TMEM) while the other can still issue / receive MMAs, commits or arrives
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ LSQ-quantizing epilogue
// ofq_gemm_lsq: the int8 GEMM of a linear layer whose output goes straight into an LSQ quantizer (qkx of the query-key
// reparameterisation, attention.py:200-207). Mainloop, ring and accumulator hand-off as gemm_tc_kernel<0, BN, STAGES>; the
// epilogue forms y = acc * rs * cs + rt * ct with the SAME roundings, then the quantizer: codes (int8), their exact 16-bit copy
// and the fp16 backward residual leave through three TMA stores per 32 x 32 chunk, the fp32 product is never written.
struct LsqEpi {
    const float* b4; const float* s_eff; const float* inv_s; const float* dot_u;
    float* part; int ld_part;          // [M, ceil(N / 32)] per-chunk partial row dots (summed in fixed order afterwards)
    int nseg, seg_len;
    FastDiv fd_period, fd_seglen;
    float qlo, qhi;
    int has16, has_res, f16;
};

template <int BN, int STAGES>
struct LsqSmem {
    static constexpr uint32_t A_BYTES = BM * KBYTES;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + BN * KBYTES;
    static constexpr uint32_t CHUNK_BYTES = 5120;                      // codes 1 KB (32B swizzle) | codes16 2 KB | res16 2 KB (64B swizzle)
    static constexpr uint32_t OUT_OFF = STAGES * STAGE_BYTES;
    static constexpr uint32_t OUT_BYTES = EPI_WARPS * 2 * CHUNK_BYTES;
    static constexpr uint32_t VEC_OFF = OUT_OFF + OUT_BYTES;           // per warp and owned chunk: cs | ct | b4 | u (32 floats each)
    static constexpr uint32_t BAR_OFF = VEC_OFF + EPI_WARPS * epi_nch(BN) * 128 * 4;
    static constexpr uint32_t TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr size_t DYN_BYTES = TOTAL + 1024;
};

__device__ __forceinline__ void st_shared_v4_b32(void* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(smem_u32(dst)), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ uint32_t cvt_pack16(float lo, float hi, bool f16) {
    uint32_t r;
    if (f16) asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    else     asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_pin8(uint32_t (&r)[8]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) : : "memory");
}

// Eight columns (piece `sub` of a 32 x 32 chunk) of row `lane`: y, quantizer, outputs into the swizzled staging rows.
// MODE 0: codes only; 1: + 16-bit copy; 2: + fp16 residual. Returns sum_j u[j] * q[j] over the piece.
// Kept small and called from a rolled loop: a fully unrolled chunk (x 2 register buffers x variants) did not fit the
// instruction cache and ran at a third of this speed.
template <int MODE, bool RT1>
__device__ __forceinline__ float lsq_piece(const uint32_t (&rr)[8], const float rsv, const float rtv, const float* __restrict__ vec,
                                           const float se, const float inv, const float qlo, const float qhi, const bool f16,
                                           uint8_t* __restrict__ row8, uint8_t* __restrict__ row16, uint8_t* __restrict__ rowr,
                                           const int sw8, const int sw16, const int sub) {
    constexpr float MAGIC = 12582912.f;                 // 1.5 * 2^23: (t + MAGIC) - MAGIC = rint(t) for |t| < 2^22; low byte of the sum's bits = the code
    const float* vp = vec + 8 * sub;
    const float4 csA = ld_shared_v4_nc(vp), csB = ld_shared_v4_nc(vp + 4);
    const float4 ctA = ld_shared_v4_nc(vp + 32), ctB = ld_shared_v4_nc(vp + 36);
    const float4 b4A = ld_shared_v4_nc(vp + 64), b4B = ld_shared_v4_nc(vp + 68);
    const float4 uA = ld_shared_v4_nc(vp + 96), uB = ld_shared_v4_nc(vp + 100);
    const float csv[8] = {csA.x, csA.y, csA.z, csA.w, csB.x, csB.y, csB.z, csB.w};
    const float ctv[8] = {ctA.x, ctA.y, ctA.z, ctA.w, ctB.x, ctB.y, ctB.z, ctB.w};
    const float b4v[8] = {b4A.x, b4A.y, b4A.z, b4A.w, b4B.x, b4B.y, b4B.z, b4B.w};
    const float uv[8] = {uA.x, uA.y, uA.z, uA.w, uB.x, uB.y, uB.z, uB.w};
    // Packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2: two columns per instruction, the same roundings per element): this
    // epilogue is bound by instruction issue, not by memory. The accumulators are small integers (|acc| < 2^22): int -> float
    // through the magic constant (integer add + FADD2) instead of I2F.
    const float2 rs2 = make_float2(rsv, rsv), rt2 = make_float2(rtv, rtv), inv2 = make_float2(inv, inv);
    const float2 magic2 = make_float2(MAGIC, MAGIC), nmagic2 = make_float2(-MAGIC, -MAGIC);
    float2 a2[4], v2[4], t2[4], mg2[4], r2[4];
    bool redo = false;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const float2 acc = __fadd2_rn(make_float2(__uint_as_float(rr[2 * h] + 0x4B400000u), __uint_as_float(rr[2 * h + 1] + 0x4B400000u)), nmagic2);
        const float2 t0 = __fmul2_rn(acc, rs2);
        const float2 y = __ffma2_rn(t0, make_float2(csv[2 * h], csv[2 * h + 1]),
                                    RT1 ? make_float2(ctv[2 * h], ctv[2 * h + 1]) : __fmul2_rn(rt2, make_float2(ctv[2 * h], ctv[2 * h + 1])));
        a2[h] = __fadd2_rn(y, make_float2(b4v[2 * h], b4v[2 * h + 1]));
        v2[h] = __fmul2_rn(a2[h], inv2);
        t2[h] = make_float2(fminf(fmaxf(v2[h].x, qlo), qhi), fminf(fmaxf(v2[h].y, qlo), qhi));
        mg2[h] = __fadd2_rn(t2[h], magic2);
        r2[h] = __fadd2_rn(mg2[h], nmagic2);
        const float2 d = __fadd2_rn(t2[h], make_float2(-r2[h].x, -r2[h].y));
        redo |= fmaxf(fabsf(d.x), fabsf(d.y)) > 0.4998f;
    }
    if (redo) {                                         // a quotient near a rounding boundary: IEEE division for the piece (rare)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            r2[h] = make_float2(rintf(fminf(fmaxf(__fdiv_rn(a2[h].x, se), qlo), qhi)), rintf(fminf(fmaxf(__fdiv_rn(a2[h].y, se), qlo), qhi)));
            mg2[h] = __fadd2_rn(r2[h], magic2);
        }
    }
    float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 4; ++h) dot2 = __ffma2_rn(make_float2(uv[2 * h], uv[2 * h + 1]), r2[h], dot2);
    const float dot = dot2.x + dot2.y;
    // eight int8 codes: the low bytes of the magic sums
    uint32_t c8[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t lo2 = __byte_perm(__float_as_uint(mg2[2 * h].x), __float_as_uint(mg2[2 * h].y), 0x0040);
        const uint32_t hi2 = __byte_perm(__float_as_uint(mg2[2 * h + 1].x), __float_as_uint(mg2[2 * h + 1].y), 0x0040);
        c8[h] = __byte_perm(lo2, hi2, 0x5410);
    }
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};\n" ::"r"(smem_u32(row8 + (((sub >> 1) ^ sw8) << 4) + ((sub & 1) << 3))), "r"(c8[0]), "r"(c8[1]));
    if (MODE >= 1)
        st_shared_v4_b32(row16 + ((sub ^ sw16) << 4), cvt_pack16(r2[0].x, r2[0].y, f16), cvt_pack16(r2[1].x, r2[1].y, f16),
                         cvt_pack16(r2[2].x, r2[2].y, f16), cvt_pack16(r2[3].x, r2[3].y, f16));
    if (MODE >= 2) {
        // inside the clamp range <=> the clamp changed nothing; outside, the sign of v tells the side (qlo <= 0 <= qhi)
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float2 dv = __fadd2_rn(r2[h], make_float2(-v2[h].x, -v2[h].y));
            const float rx = t2[h].x == v2[h].x ? dv.x : copysignf(2.f, v2[h].x);
            const float ry = t2[h].y == v2[h].y ? dv.y : copysignf(2.f, v2[h].y);
            w[h] = cvt_pack16(rx, ry, true);
        }
        st_shared_v4_b32(rowr + ((sub ^ sw16) << 4), w[0], w[1], w[2], w[3]);
    }
    return dot;
}

// sixteen columns = two 8-column pieces
template <int MODE, bool RT1>
__device__ __forceinline__ float lsq_piece16(const uint32_t (&rr)[16], const float rsv, const float rtv, const float* __restrict__ vec,
                                             const float se, const float inv, const float qlo, const float qhi, const bool f16,
                                             uint8_t* __restrict__ row8, uint8_t* __restrict__ row16, uint8_t* __restrict__ rowr,
                                             const int sw8, const int sw16, const int sub) {
    const uint32_t lo[8] = {rr[0], rr[1], rr[2], rr[3], rr[4], rr[5], rr[6], rr[7]};
    const uint32_t hi[8] = {rr[8], rr[9], rr[10], rr[11], rr[12], rr[13], rr[14], rr[15]};
    const float d0 = lsq_piece<MODE, RT1>(lo, rsv, rtv, vec, se, inv, qlo, qhi, f16, row8, row16, rowr, sw8, sw16, sub);
    const float d1 = lsq_piece<MODE, RT1>(hi, rsv, rtv, vec, se, inv, qlo, qhi, f16, row8, row16, rowr, sw8, sw16, sub + 1);
    return d0 + d1;
}

template <int BN, int STAGES, int MODE, bool RT1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_lsq_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmQ8, const __grid_constant__ CUtensorMap tmQ16,
                const __grid_constant__ CUtensorMap tmR16, const GemmParams p, const LsqEpi e, const int num_tiles,
                const int mtiles, const int ntiles) {
    using L = LsqSmem<BN, STAGES>;
    constexpr uint32_t A_BYTES = L::A_BYTES;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES;
    constexpr uint32_t ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    constexpr uint32_t UMMA_K_BYTES = 32;
    const uint32_t IDESC = umma_idesc(2u, 1u, BM, BN);

    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* vec_s = reinterpret_cast<float*>(smem + L::VEC_OFF);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmQ8);
        tma_prefetch_desc(&tmQ16); tma_prefetch_desc(&tmR16);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS); }
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

    if (warp == 0) {
        if (elect_one()) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_5d(sa, &tmA, &full_bar[s], (c.it_begin + i) * KBYTES, c.m0, 0, 0, 0);
                    tma_load_5d(sa + A_BYTES, &tmB, &full_bar[s], (c.it_begin + i) * KBYTES, c.n0, 0, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            uint32_t it = 0, tc = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_kmajor_sw128(sa), bdesc = umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
                    for (uint32_t kk = 0; kk < KBYTES / UMMA_K_BYTES; ++kk)
                        umma_i8(tmem_d, adesc + kk * (UMMA_K_BYTES >> 4), bdesc + kk * (UMMA_K_BYTES >> 4), IDESC, (i | kk) != 0);
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&acc_full[as]);
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int NCH = epi_nch(BN);
        uint8_t* stage_base = smem + L::OUT_OFF + (warp - 2) * 2 * L::CHUNK_BYTES;
        float* myvec = vec_s + (warp - 2) * NCH * 128;
        uint32_t tc = 0, chunk = 0;
        const float qlo = e.qlo, qhi = e.qhi;
        const bool f16 = e.f16 != 0;
        // vectors of the NEXT tile, loaded one tile ahead: column `lane` of every owned chunk, the row's rs / rt and the step
        // sizes of the (at most two: BN <= seg_len) segments the tile touches
        float n_cs[NCH], n_ct[NCH], n_b4[NCH], n_u[NCH], n_rs = 0.f, n_rt = 0.f, n_se[2] = {1.f, 1.f}, n_inv[2] = {1.f, 1.f};
        TileCoord cn;
        auto load_vecs = [&](const TileCoord& c) {
            const int m = c.m0 + q * 32 + lane;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int n = c.n0 + (half + 2 * k) * 32 + lane;
                const bool ok = n < p.N;
                n_cs[k] = ok ? (p.cs.p ? __ldg(p.cs.p + p.cs.fd.mod(n)) : 1.0f) : 0.f;
                n_ct[k] = (ok && p.has_rank1) ? (p.ct.p ? __ldg(p.ct.p + n) : 1.0f) : 0.f;
                n_b4[k] = (ok && e.b4) ? __ldg(e.b4 + n) : 0.f;
                n_u[k] = (ok && e.dot_u) ? __ldg(e.dot_u + n) : 0.f;
            }
            const bool row_ok = m < p.M;
            n_rs = row_ok ? (p.rs.p ? __ldg(p.rs.p + p.rs.fd.mod(m)) : 1.0f) : 0.f;
            n_rt = (row_ok && p.has_rank1) ? (p.rt.p ? __ldg(p.rt.p + p.rt.fd.mod(m)) : 1.0f) : 0.f;
            const int seg0 = e.fd_seglen.div(c.n0);
            const int mp = e.fd_period.mod(row_ok ? m : 0);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int sg = min(seg0 + k, e.nseg - 1);
                n_se[k] = __ldg(e.s_eff + mp * e.nseg + sg);
                n_inv[k] = __ldg(e.inv_s + mp * e.nseg + sg);
            }
        };
        if ((int)blockIdx.x < num_tiles) {
            cn = decode_tile(p, blockIdx.x, BN, mtiles, ntiles);
            load_vecs(cn);
        }
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
            const TileCoord c = cn;
            const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
            const float rsv = n_rs, rtv = n_rt;
            const float se_a = n_se[0], se_b = n_se[1], inv_a = n_inv[0], inv_b = n_inv[1];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                myvec[k * 128 + lane] = n_cs[k];
                myvec[k * 128 + 32 + lane] = n_ct[k];
                myvec[k * 128 + 64 + lane] = n_b4[k];
                myvec[k * 128 + 96 + lane] = n_u[k];
            }
            __syncwarp();
            if (t + (int)gridDim.x < num_tiles) {
                cn = decode_tile(p, t + gridDim.x, BN, mtiles, ntiles);
                load_vecs(cn);
            }
            if (c.nit > 0) {
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
            }
            const uint32_t tmem_acc = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            int nvalid = (min(BN, p.N - c.n0) + 31) / 32;
            const int m_row0 = c.m0 + q * 32;
            if (m_row0 >= p.M) nvalid = 0;
            const int seg_first = e.fd_seglen.div(c.n0);
            const int seg_switch = (seg_first + 1) * e.seg_len;          // first column of the tile's second segment
            // this warp's chunks (half, half + 2, ...), each as two 16-column pieces; the TMEM loads ping-pong between two
            // 16-register buffers one piece ahead (also across chunk boundaries). ROLLED loop over the chunks: a fully unrolled
            // tile did not fit the instruction cache.
            const int nown = nvalid > half ? (nvalid - half + 1) >> 1 : 0;
            uint32_t ra[16], rb[16];
            if (nown > 0) tmem_ld_32x32_x16r(tmem_acc + (uint32_t)(half * 32), ra);
            const int sw8 = (lane >> 2) & 1, sw16 = (lane >> 1) & 3;
#pragma unroll 1
            for (int k = 0; k < nown; ++k) {
                const int cc = half + 2 * k, n_c = c.n0 + cc * 32;
                const float* vec = myvec + k * 128;
                uint8_t* buf = stage_base + (chunk & 1) * L::CHUNK_BYTES;
                if (chunk >= 2) {              // the staging buffer used two chunks ago must have been read by TMA
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                }
                ++chunk;
                const bool second = n_c >= seg_switch;
                const float se = second ? se_b : se_a, inv = second ? inv_b : inv_a;
                uint8_t* row8 = buf + lane * 32;
                uint8_t* row16 = buf + 1024 + lane * 64;
                uint8_t* rowr = buf + 3072 + lane * 64;
                float dot = 0.f;
                tmem_ld_wait(); tmem_ld_pin16(ra);
                tmem_ld_32x32_x16r(tmem_acc + (uint32_t)(cc * 32 + 16), rb);
                if (!(p.debug_nostore & 2)) dot += lsq_piece16<MODE, RT1>(ra, rsv, rtv, vec, se, inv, qlo, qhi, f16, row8, row16, rowr, sw8, sw16, 0);
                tmem_ld_wait(); tmem_ld_pin16(rb);
                if (k + 1 < nown) tmem_ld_32x32_x16r(tmem_acc + (uint32_t)((cc + 2) * 32), ra);
                if (!(p.debug_nostore & 2)) dot += lsq_piece16<MODE, RT1>(rb, rsv, rtv, vec, se, inv, qlo, qhi, f16, row8, row16, rowr, sw8, sw16, 2);
                if (e.part && m_row0 + lane < p.M && !(p.debug_nostore & 8)) e.part[(long long)(m_row0 + lane) * e.ld_part + (n_c >> 5)] = dot;
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(p.debug_nostore & 1)) {
                    tma_store_5d(&tmQ8, buf, n_c, m_row0, 0, 0, 0);
                    if (MODE >= 1) tma_store_5d(&tmQ16, buf + 1024, n_c, m_row0, 0, 0, 0);
                    if (MODE >= 2) tma_store_5d(&tmR16, buf + 3072, n_c, m_row0, 0, 0, 0);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ dX GEMM + LSQ backward epilogue
// ofq_gemm_dx_lsq: dX_hat = A16 . W16 (the backward GEMM of a quantized linear layer, qlinear.py:58-73) whose epilogue IS the
// backward of the layer's input quantizer (autograd of lsq.py:571-602): the x tile arrives by TMA beside the accumulator, the
// epilogue applies the straight-through mask, stores dx, and leaves per-row partials of the step-size gradient and per-warp
// column sums (d move_b4, d move_aft) for the ordinary finalize pass. dX_hat never exists in HBM.
struct DxLsqEpi {
    unsigned int* amax;     // optional: max |dx| over the whole problem (bits of a non-negative float, atomicMax; pre-set to 0)
    const float* b4;        // [N]
    float* rowpart;         // [2 * ntiles][M]
    float* colpart;         // [4 * mtiles][3][N]
    float qlo, qhi;
};

template <int BN, int STAGES>
struct DxLsqSmem {
    static constexpr uint32_t A_BYTES = BM * KBYTES;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + BN * KBYTES;
    static constexpr uint32_t WARP_BYTES = 3 * 4096;                  // x tile (double buffered) | dx staging, 32 x 128 B each, 128B swizzle
    static constexpr uint32_t OUT_OFF = STAGES * STAGE_BYTES;
    static constexpr uint32_t VEC_OFF = OUT_OFF + EPI_WARPS * WARP_BYTES;           // per warp and owned chunk: b4[32]
    static constexpr uint32_t BAR_OFF = VEC_OFF + EPI_WARPS * epi_nch(BN) * 32 * 4;
    static constexpr uint32_t TOTAL = BAR_OFF + (2 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
    static constexpr size_t DYN_BYTES = TOTAL + 1024;
};

// sixteen columns (half `sub` of a 32 x 32 chunk) of row `lane`; returns the row's partial of dy * (q - v | q)
__device__ __forceinline__ float dxlsq_piece(const uint32_t (&rr)[16], const float rsv, const float csv, const float* __restrict__ b4v,
                                             const uint8_t* __restrict__ xrow, uint8_t* __restrict__ orow, const float qlo, const float qhi,
                                             const int lane, const int sub, float& omax) {
    constexpr float MAGIC = 12582912.f;
    const float2 rs2 = make_float2(rsv, rsv), cs2 = make_float2(csv, csv);
    float2 part2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int j4 = sub * 4 + g;                                   // 16-byte group of the 128-byte row
        const float4 x4 = ld_shared_v4_nc(xrow + ((j4 ^ (lane & 7)) << 4));
        const float4 b44 = ld_shared_v4_nc(b4v + sub * 16 + 4 * g);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, bs[4] = {b44.x, b44.y, b44.z, b44.w};
        float o[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 4 * g + 2 * h;
            const float2 y = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), rs2), cs2);
            const float2 v = __fmul2_rn(__fadd2_rn(make_float2(xs[2 * h], xs[2 * h + 1]), make_float2(bs[2 * h], bs[2 * h + 1])), rs2);
            const float2 t = make_float2(fminf(fmaxf(v.x, qlo), qhi), fminf(fmaxf(v.y, qlo), qhi));
            const float2 q = __fadd2_rn(__fadd2_rn(t, make_float2(MAGIC, MAGIC)), make_float2(-MAGIC, -MAGIC));
            const float2 qv = __fadd2_rn(q, make_float2(-v.x, -v.y));
            const bool in0 = t.x == v.x, in1 = t.y == v.y;            // inside the clamp range <=> the clamp changed nothing
            part2 = __ffma2_rn(y, make_float2(in0 ? qv.x : q.x, in1 ? qv.y : q.y), part2);
            o[2 * h] = in0 ? y.x : 0.f;
            o[2 * h + 1] = in1 ? y.y : 0.f;
            omax = fmaxf(omax, fmaxf(fabsf(o[2 * h]), fabsf(o[2 * h + 1])));
        }
        st_shared_v4_nc(orow + ((j4 ^ (lane & 7)) << 4), o[0], o[1], o[2], o[3]);
    }
    return part2.x + part2.y;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_dxlsq_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmC, const GemmParams p,
                  const DxLsqEpi e, const int num_tiles, const int mtiles, const int ntiles) {
    using L = DxLsqSmem<BN, STAGES>;
    constexpr uint32_t A_BYTES = L::A_BYTES;
    constexpr uint32_t STAGE_BYTES = L::STAGE_BYTES;
    constexpr uint32_t ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
    constexpr uint32_t UMMA_K_BYTES = 32;
    const uint32_t IDESC = umma_idesc(1u, (uint32_t)p.ab_fmt, BM, BN) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16);

    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* x_bar = acc_empty + 2;             // [EPI_WARPS][2]: x tile of a chunk has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + 2 * EPI_WARPS);
    float* vec_s = reinterpret_cast<float*>(smem + L::VEC_OFF);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmC);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS); }
        for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&x_bar[s], 1);
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

    if (warp == 0) {
        if (elect_one()) {
            constexpr int kelem = KBYTES / 2;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int kb = c.it_begin + i;
                    uint8_t* sa = smem + s * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    if (p.a_mn) {
                        for (int j = 0; j < BM / 64; ++j) tma_load_5d(sa + j * 8192, &tmA, &full_bar[s], c.m0 + 64 * j, kb * kelem, 0, 0, 0);
                    } else {
                        tma_load_5d(sa, &tmA, &full_bar[s], kb * kelem, c.m0, 0, 0, 0);
                    }
                    if (p.b_mn) {
                        for (int j = 0; j < BN / 64; ++j) tma_load_5d(sb + j * 8192, &tmB, &full_bar[s], c.n0 + 64 * j, kb * kelem, 0, 0, 0);
                    } else {
                        tma_load_5d(sb, &tmB, &full_bar[s], kb * kelem, c.n0, 0, 0, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            uint32_t it = 0, tc = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
                const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
                const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                for (int i = 0; i < c.nit; ++i, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const bool bmn = p.b_mn != 0, amn = p.a_mn != 0;
                    const uint64_t bdesc = bmn ? umma_desc_mnmajor_sw128(sa + A_BYTES) : umma_desc_kmajor_sw128(sa + A_BYTES);
                    const uint64_t adesc = amn ? umma_desc_mnmajor_sw128(sa) : umma_desc_kmajor_sw128(sa);
                    const uint64_t badv = bmn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
                    const uint64_t aadv = amn ? (2048u >> 4) : (UMMA_K_BYTES >> 4);
#pragma unroll
                    for (uint32_t kk = 0; kk < KBYTES / UMMA_K_BYTES; ++kk)
                        umma_f16(tmem_d, adesc + kk * aadv, bdesc + kk * badv, IDESC, (i | kk) != 0);
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&acc_full[as]);
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int NCH = epi_nch(BN);
        uint8_t* wbase = smem + L::OUT_OFF + (warp - 2) * L::WARP_BYTES;
        uint8_t* obuf = wbase + 8192;
        uint64_t* xb = x_bar + 2 * (warp - 2);
        float* myvec = vec_s + (warp - 2) * NCH * 32;
        const float qlo = e.qlo, qhi = e.qhi;
        const float csv = p.cs.p ? __ldg(p.cs.p) : 1.0f;              // the range un-scale: one value for the whole problem
        uint32_t tc = 0, nstore = 0, xdone = 0, xissued = 0;
        float omax = 0.f;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tc) {
            const TileCoord c = decode_tile(p, t, BN, mtiles, ntiles);
            const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
            const int m_row0 = c.m0 + q * 32, m = m_row0 + lane;
            const bool row_ok = m < p.M;
            const float rsv = row_ok ? (p.rs.p ? __ldg(p.rs.p + p.rs.fd.mod(m)) : 1.0f) : 0.f;
            const int nvalid = (min(BN, p.N - c.n0) + 31) / 32;
            const int nown = nvalid > half ? (nvalid - half + 1) >> 1 : 0;
            // first x tile of this tile's chunks, and the b4 vectors of all owned chunks
            if (nown > 0 && lane == 0) {
                uint64_t* bar = &xb[xissued & 1];
                mbar_arrive_expect_tx(bar, 4096);
                tma_load_5d(wbase + (xissued & 1) * 4096, &tmX, bar, c.n0 + half * 32, m_row0, 0, 0, 0);
            }
            if (nown > 0) ++xissued;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int n = c.n0 + (half + 2 * k) * 32 + lane;
                myvec[k * 32 + lane] = (n < p.N && e.b4) ? __ldg(e.b4 + n) : 0.f;
            }
            __syncwarp();
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t ra[16], rb[16];
            if (nown > 0) tmem_ld_32x32_x16r(tmem_acc + (uint32_t)(half * 32), ra);
            float part = 0.f;
#pragma unroll 1
            for (int k = 0; k < nown; ++k) {
                const int cc = half + 2 * k, n_c = c.n0 + cc * 32;
                if (k + 1 < nown) {                  // x tile of the next chunk into the other buffer
                    if (lane == 0) {
                        uint64_t* bar = &xb[xissued & 1];
                        mbar_arrive_expect_tx(bar, 4096);
                        tma_load_5d(wbase + (xissued & 1) * 4096, &tmX, bar, n_c + 64, m_row0, 0, 0, 0);
                    }
                    ++xissued;
                }
                const uint32_t xslot = xdone & 1u;               // loads are consumed in issue order, alternating buffers
                mbar_wait(&xb[xslot], (xdone >> 1) & 1u);
                ++xdone;
                if (nstore > 0) {                    // the staging buffer must have been read by the previous chunk's store
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                }
                const uint8_t* xrow = wbase + xslot * 4096 + lane * 128;
                uint8_t* orow = obuf + lane * 128;
                const float* b4v = myvec + k * 32;
                tmem_ld_wait(); tmem_ld_pin16(ra);
                tmem_ld_32x32_x16r(tmem_acc + (uint32_t)(cc * 32 + 16), rb);
                part += dxlsq_piece(ra, rsv, csv, b4v, xrow, orow, qlo, qhi, lane, 0, omax);
                tmem_ld_wait(); tmem_ld_pin16(rb);
                if (k + 1 < nown) tmem_ld_32x32_x16r(tmem_acc + (uint32_t)((cc + 2) * 32), ra);
                part += dxlsq_piece(rb, rsv, csv, b4v, xrow, orow, qlo, qhi, lane, 1, omax);
                fence_proxy_async_smem();
                __syncwarp();
                // d move_b4: column sums of the staged dx tile over this warp's 32 rows. Lane l walks column l down the rows
                // (row r holds it in 16-byte group (l / 4) ^ (r % 8): one conflict-free 128-byte wavefront per row). d move_aft
                // (the unmasked column sums) is not formed here at all: sum_m dX_hat = (colsum(dY) * colscale) . W_codes.
                {
                    float cb = 0.f;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r)
                        cb += *reinterpret_cast<const float*>(obuf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                    const int col = n_c + lane;
                    if (col < p.N) e.colpart[((long long)(c.m0 / BM) * 4 + q) * 3 * p.N + p.N + col] = cb;
                }
                if (lane == 0) {
                    tma_store_5d(&tmC, obuf, n_c, m_row0, 0, 0, 0);
                    tma_store_commit();
                }
                ++nstore;
            }
            if (row_ok && nown > 0) e.rowpart[(long long)((c.n0 / BN) * 2 + half) * p.M + m] = part;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (e.amax) {
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o2));
            if (lane == 0) atomicMax(e.amax, __float_as_uint(omax));
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// d move_aft of a linear layer's input from colsum(dY) and the weight codes: sum_n u1[n] * u2[n] * codes[n][k], spread over the
// `nslots` column-partial slots the finalize pass sums anyway (slot b takes rows b, b + nslots, ...; vector 0 of colpart)
__global__ void __launch_bounds__(128)
codes_vecmat_parts_kernel(const int8_t* __restrict__ codes, int rows, int cols, long long ld, const float* __restrict__ u1,
                          const float* __restrict__ u2, float* __restrict__ colpart, int nslots) {
    const int b = blockIdx.x;
    for (int col = threadIdx.x; col < cols; col += blockDim.x) {
        float acc = 0.f;
        for (int n = b; n < rows; n += nslots) acc = fmaf(__ldg(u1 + n) * (u2 ? __ldg(u2 + n) : 1.f), (float)codes[(long long)n * ld + col], acc);
        colpart[(long long)b * 3 * cols + col] = acc;
    }
}

// rowdot[m, s] = sum of the per-chunk partials of segment s, in column order (deterministic)
__global__ void __launch_bounds__(256)
rowdot_reduce_kernel(const float* __restrict__ part, int ld_part, long long M, int nseg, int chunks_per_seg, float* __restrict__ rowdot) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * nseg) return;
    const long long m = i / nseg;
    const int sg = (int)(i - m * nseg);
    const float* pp = part + m * ld_part + (long long)sg * chunks_per_seg;
    float a = 0.f;
    for (int j = 0; j < chunks_per_seg; ++j) a += pp[j];
    rowdot[i] = a;
}

// ------------------------------------------------------------------------------------------ host side
template <int KIND, int BN, int STAGES, int NA = 1, int OUT_BUFS = 2>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       GemmParams p, cudaStream_t stream) {
    constexpr size_t smem = SmemLayout<BN, STAGES, NA, OUT_BUFS>::DYN_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_tc_kernel<KIND, BN, STAGES, NA, OUT_BUFS>;
    static bool configured = false;  // per instantiation; attribute is per function, idempotent
    if (!configured) {
        OFQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int mtiles = (p.M + BM - 1) / BM, ntiles = (p.N + BN - 1) / BN;
    p.fd_ntiles = make_fastdiv(ntiles); p.fd_mtiles = make_fastdiv(mtiles);
    const long long tiles = (long long)mtiles * ntiles * p.nb1 * p.nb2 * p.splits;
    if (tiles > 0x7fffffff) {
        ofq_set_error("ofq_gemm: too many tiles");
        return OFQ_ERR_ARG;
    }
    const int grid = (int)(tiles < ofq_num_sms() ? tiles : ofq_num_sms());   // one persistent CTA per SM
    if constexpr (NA == 1 && BN >= 64) {
        // 16-epilogue-warp variant (OFQ_GEMM_EPI16: 1 = on for every single-CTA launch, 0 = off)
        static const int epi16 = [] { const char* e = getenv("OFQ_GEMM_EPI16"); return e ? atoi(e) : OFQ_GEMM_EPI16_DEFAULT; }();
        if (epi16 && !p.debug_nostore) {
            constexpr int ST16 = BN >= 192 ? (STAGES > 3 ? 3 : STAGES) : (STAGES > 4 ? 4 : STAGES);     // 64 KB of staging beside the ring
            constexpr size_t smem16 = SmemLayout16<BN, ST16>::DYN_BYTES;
            static_assert(smem16 <= 227 * 1024, "shared memory budget exceeded");
            auto kern16 = gemm_tc16_kernel<KIND, BN, ST16>;
            static bool configured16 = false;
            if (!configured16) {
                OFQ_CUDA(cudaFuncSetAttribute(kern16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
                configured16 = true;
            }
            kern16<<<grid, NUM_THREADS16, smem16, stream>>>(tmA, tmB, tmC, p, (int)tiles, mtiles, ntiles);
            OFQ_CUDA(cudaGetLastError());
            return 0;
        }
    }
    kern<<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, tmC, p, (int)tiles, mtiles, ntiles);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}


// Largest number of co-resident CTA pairs for an instantiation (clusters need two free SMs of one TPC).
template <int KIND, int BN, int STAGES, int OUT_BUFS = 2>
static int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                            GemmParams p, cudaStream_t stream) {
    constexpr size_t smem = SmemLayout2<BN, STAGES, OUT_BUFS>::DYN_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_tc_pair_kernel<KIND, BN, STAGES, OUT_BUFS>;
    static int max_pairs = 0;         // per instantiation
    if (max_pairs == 0) {
        OFQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(ofq_num_sms() & ~1);
        q.blockDim = dim3(NUM_THREADS);
        q.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.attrs = at; q.numAttrs = 1;
        int n = 0;
        OFQ_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
        if (n <= 0) { ofq_set_error("ofq_gemm: no CTA pair fits on this device"); return OFQ_ERR_CUDA; }
        max_pairs = n < ofq_num_sms() / 2 ? n : ofq_num_sms() / 2;
    }
    const int mtiles = (p.M + 2 * BM - 1) / (2 * BM), ntiles = (p.N + BN - 1) / BN;
    p.fd_ntiles = make_fastdiv(ntiles); p.fd_mtiles = make_fastdiv(mtiles);
    const long long tiles = (long long)mtiles * ntiles * p.nb1 * p.nb2 * p.splits;
    if (tiles > 0x7fffffff) {
        ofq_set_error("ofq_gemm: too many tiles");
        return OFQ_ERR_ARG;
    }
    const int pairs = (int)(tiles < max_pairs ? tiles : max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    OFQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p, (int)tiles, mtiles, ntiles));
    return 0;
}

static VecRef make_vec(const ofq_vec_t* v) {
    VecRef r;
    r.p = v ? v->ptr : nullptr;
    r.period = (v && v->period > 0) ? v->period : 0x7fffffff;
    r.fd = make_fastdiv(r.period);
    r.bs1 = v ? v->bstride1 : 0;
    r.bs2 = v ? v->bstride2 : 0;
    return r;
}

}  // namespace ofq

using namespace ofq;

// Build a 5-D tensor map {K, rows, k2, b1, b2} with a {128 bytes of K, box_rows, 1, 1, 1} box, 128B swizzle.
static int make_operand_map(CUtensorMap* tm, const ofq_operand_t* op, int elem_bytes, int rows, int K,
                            int k2, int nb1, int nb2, int box_rows, bool f16 = false) {
    int k2_eff = op->k2_stride ? (op->k2_mod > 0 ? (op->k2_mod < k2 ? op->k2_mod : k2) : k2) : 1;
    if (op->dual_delta > 0) k2_eff = k2 + op->dual_delta;     // slices [0, k2) and [delta, delta + k2) are both addressed
    cuuint64_t dims[5] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)k2_eff,
                          (cuuint64_t)(op->bstride1 ? nb1 : 1), (cuuint64_t)(op->bstride2 ? nb2 : 1)};
    const cuuint64_t row_b = (cuuint64_t)op->row_stride * elem_bytes;
    cuuint64_t strides[4] = {row_b,
                             (cuuint64_t)(op->k2_stride ? op->k2_stride * elem_bytes : row_b),
                             (cuuint64_t)(op->bstride1 ? op->bstride1 * elem_bytes : row_b),
                             (cuuint64_t)(op->bstride2 ? op->bstride2 * elem_bytes : row_b)};
    for (int i = 0; i < 4; ++i)
        if (strides[i] % 16) {
            ofq_set_error("gemm operand stride %d (%llu bytes) is not a multiple of 16", i,
                          (unsigned long long)strides[i]);
            return OFQ_ERR_ARG;
        }
    if (reinterpret_cast<uintptr_t>(op->ptr) % 16) {
        ofq_set_error("gemm operand pointer is not 16-byte aligned");
        return OFQ_ERR_ARG;
    }
    cuuint32_t box[5] = {(cuuint32_t)(KBYTES / elem_bytes), (cuuint32_t)box_rows, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (op->mn_major) {   // {rows (contiguous), K, k2, b1, b2}; box = 64 rows (128 B) x 64 k
        dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
        box[0] = 64; box[1] = (cuuint32_t)(KBYTES / elem_bytes);
    }
    return ofq_encode_tensor_map(tm, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16),
                                 5, const_cast<void*>(op->ptr), dims, strides, box, estr);
}

// Output map {N, M, 1, b1, b2}, fp32, 32 x 32 box (one epilogue warp's chunk), 128B swizzle.
static int make_out_map(CUtensorMap* tm, const ofq_gemm_out_t* out, int M, int N, int nb1, int nb2) {
    if (reinterpret_cast<uintptr_t>(out->ptr) % 16 || out->ld % 4 || out->bstride1 % 4 || out->bstride2 % 4) {
        ofq_set_error("gemm output must be 16-byte aligned with ld / batch strides that are multiples of 4 floats");
        return OFQ_ERR_ARG;
    }
    const cuuint64_t row_b = (cuuint64_t)out->ld * 4;
    cuuint64_t dims[5] = {(cuuint64_t)N, (cuuint64_t)M, 1, (cuuint64_t)(out->bstride1 ? nb1 : 1),
                          (cuuint64_t)(out->bstride2 ? nb2 : 1)};
    cuuint64_t strides[4] = {row_b, row_b, (cuuint64_t)(out->bstride1 ? out->bstride1 * 4 : row_b),
                             (cuuint64_t)(out->bstride2 ? out->bstride2 * 4 : row_b)};
    cuuint32_t box[5] = {32, 32, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return ofq_encode_tensor_map(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, out->ptr, dims, strides, box, estr);
}

extern "C" int ofq_gemm(int kind, const ofq_operand_t* A, const ofq_operand_t* B, const ofq_gemm_out_t* out,
                        int M, int N, int K, int k2, int nb1, int nb2, int splits, const ofq_vec_t* rs,
                        const ofq_vec_t* cs, const ofq_vec_t* rt, const ofq_vec_t* ct, void* stream) {
    return ofq_gemm_ex(kind, A, B, out, M, N, K, k2, nb1, nb2, splits, rs, cs, rt, ct, nullptr, stream);
}

extern "C" int ofq_gemm_ex(int kind, const ofq_operand_t* A, const ofq_operand_t* B, const ofq_gemm_out_t* out,
                           int M, int N, int K, int k2, int nb1, int nb2, int splits, const ofq_vec_t* rs,
                           const ofq_vec_t* cs, const ofq_vec_t* rt, const ofq_vec_t* ct, float* out_absmax, void* stream) {
    if (kind != OFQ_GEMM_I8 && kind != OFQ_GEMM_BF16 && kind != OFQ_GEMM_F16) {
        ofq_set_error("ofq_gemm: unknown kind %d", kind);
        return OFQ_ERR_ARG;
    }
    if (M <= 0 || N <= 0 || K <= 0 || k2 <= 0 || nb1 <= 0 || nb2 <= 0 || splits < 0) {
        ofq_set_error("ofq_gemm: non-positive extent");
        return OFQ_ERR_ARG;
    }
    OFQ_CHECK_ARCH();
    const int eb = kind == OFQ_GEMM_I8 ? 1 : 2;
    const int kelem = KBYTES / eb;
    GemmParams p;
    p.M = M; p.N = N;
    p.kblocks = (K + kelem - 1) / kelem;
    p.k2 = k2; p.splits = splits;
    p.nb1 = nb1; p.nb2 = nb2;
    p.a_b1 = A->bstride1 != 0; p.a_b2 = A->bstride2 != 0;
    p.b_b1 = B->bstride1 != 0; p.b_b2 = B->bstride2 != 0;
    p.a_k2 = A->k2_stride != 0; p.b_k2 = B->k2_stride != 0;
    p.a_dual_delta = A->dual_delta;
    p.a_k2mod = A->k2_mod > 0 ? A->k2_mod : 0x7fffffff;
    p.b_k2mod = B->k2_mod > 0 ? B->k2_mod : 0x7fffffff;
    p.fd_nbatch = make_fastdiv(nb1 * nb2); p.fd_nb1 = make_fastdiv(nb1);
    p.fd_kblocks = make_fastdiv(p.kblocks); p.fd_ak2mod = make_fastdiv(p.a_k2mod); p.fd_bk2mod = make_fastdiv(p.b_k2mod);
    if ((long long)p.kblocks * k2 * 65 > 0x7fffffffLL || splits > 64 || (long long)nb1 * nb2 > 0x7fffffffLL) {
        ofq_set_error("ofq_gemm: at most 64 splits; k-block count x 65 and the batch count must fit 31 bits");
        return OFQ_ERR_ARG;
    }
    p.c_b1 = out->bstride1 != 0; p.c_b2 = out->bstride2 != 0;
    p.rs = make_vec(rs); p.cs = make_vec(cs); p.rt = make_vec(rt); p.ct = make_vec(ct);
    p.has_rank1 = (rt && rt->ptr) || (ct && ct->ptr);
    p.atomic = out->accumulate;
    p.ab_fmt = kind == OFQ_GEMM_F16 ? 0 : 1;
    p.a_mn = A->mn_major != 0; p.b_mn = B->mn_major != 0;
    if ((p.a_mn || p.b_mn) && (kind == OFQ_GEMM_I8 || A->dual_delta > 0)) {
        ofq_set_error("ofq_gemm: MN-major operands are supported for the 16-bit kinds without dual-A only");
        return OFQ_ERR_ARG;
    }
    p.amax = reinterpret_cast<unsigned int*>(out_absmax);
    static const int nostore = [] { const char* e = getenv("OFQ_GEMM_NOSTORE"); return e ? atoi(e) : 0; }();
    p.debug_nostore = nostore;
    if (out_absmax && (p.atomic || splits > 1)) {
        ofq_set_error("ofq_gemm: the output maximum is tracked for plain (non-accumulating, unsplit) stores only");
        return OFQ_ERR_ARG;
    }
    if (splits > 1 && !p.atomic) {
        ofq_set_error("ofq_gemm: split-K requires an accumulating (pre-zeroed) output");
        return OFQ_ERR_ARG;
    }
    // CTA pairs (cta_group::2, 256-row tiles) for the wide-N problems with many row blocks, where the third less operand
    // traffic per CTA pays (measured on B200, tools/gemm_sweep.py and the per-site bench table: +4..14 % at N >= 1024;
    // the N = 384 layers and the 198-row attention batches are faster on the single-CTA kernel with its 4-stage ring).
    // OFQ_GEMM_PAIR=0 / 2 forces the single-CTA / the pair kernel wherever it is legal (A/B measurements, tests).
    static const int pair_mode = [] { const char* e = getenv("OFQ_GEMM_PAIR"); return e ? atoi(e) : 1; }();
    const bool pair_legal = A->dual_delta == 0 && M > BM;
    const bool pair = pair_legal && (pair_mode == 2 || (pair_mode == 1 && M >= 4 * BM && N >= 1024));
    // tile width: minimise (number of N tiles) x (per-tile fixed cost + tile width); the fixed cost (pipeline fill,
    // barrier round trips, epilogue start-up) is worth about 128 columns of MMA/epilogue work
    static const int widths[] = {256, 224, 192, 128, 64, 32};
    int bn = 32;
    long long best_cost = -1;
    for (int w : widths) {
        // MN-major B tiles are whole 64-row swizzle atoms (per CTA: half the tile in pair mode)
        if (p.b_mn && w % (pair ? 128 : 64)) continue;
        const long long nt = (N + w - 1) / w;
        const long long cost = nt * (128 + w);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; bn = w; }
    }
    if (splits == 0) {
        // auto split-K (accumulating outputs only): work items = tiles x splits are dealt round-robin to one CTA (pair) per
        // SM, so the launch takes ceil(items / slots) rounds of (k-blocks per split + a fixed per-item cost: pipeline fill,
        // epilogue, reduce-add ~ 6 k-blocks); take the split count with the cheapest total.
        splits = 1;
        if (p.atomic) {
            const long long mt = pair ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM, nt = (N + bn - 1) / bn;
            const long long T = mt * nt * nb1 * nb2;
            const long long slots = pair ? ofq_num_sms() / 2 : ofq_num_sms();
            const long long KB = (long long)p.kblocks * k2;
            static const double fixed = [] { const char* e = getenv("OFQ_GEMM_SPLIT_COST"); return e ? atof(e) : 6.0; }();
            double best = -1.0;
            for (int sp = 1; sp <= 64 && sp <= (KB / 4 > 1 ? KB / 4 : 1); ++sp) {
                const long long rounds = (T * sp + slots - 1) / slots;
                const double cost = (double)rounds * ((double)((KB + sp - 1) / sp) + fixed);
                if (best < 0 || cost < best * 0.98) { best = cost; splits = sp; }      // ties / near-ties: fewer splits
            }
        }
    }
    p.splits = splits;
    p.fd_splits = make_fastdiv(splits);
    if (out_absmax && splits > 1) {
        ofq_set_error("ofq_gemm: the output maximum is tracked for plain (non-accumulating, unsplit) stores only");
        return OFQ_ERR_ARG;
    }
    CUtensorMap tmA, tmB, tmC;
    int rc = make_operand_map(&tmA, A, eb, M, K, k2, nb1, nb2, BM, kind == OFQ_GEMM_F16);
    if (rc) return rc;
    rc = make_operand_map(&tmB, B, eb, N, K, k2, nb1, nb2, pair ? bn / 2 : bn, kind == OFQ_GEMM_F16);
    if (rc) return rc;
    rc = make_out_map(&tmC, out, M, N, nb1, nb2);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pair) {
#define OFQ_DISPATCH_PAIR(KIND)                                                    \
    switch (bn) {                                                                  \
        case 256: return launch_gemm_pair<KIND, 256, 4>(tmA, tmB, tmC, p, st);     \
        case 224: return launch_gemm_pair<KIND, 224, 5>(tmA, tmB, tmC, p, st);     \
        case 192: return launch_gemm_pair<KIND, 192, 5>(tmA, tmB, tmC, p, st);     \
        case 128: return launch_gemm_pair<KIND, 128, 6>(tmA, tmB, tmC, p, st);     \
        case 64:  return launch_gemm_pair<KIND, 64, 6>(tmA, tmB, tmC, p, st);      \
        default:  return launch_gemm_pair<KIND, 32, 6>(tmA, tmB, tmC, p, st);      \
    }
        if (kind == OFQ_GEMM_I8) { OFQ_DISPATCH_PAIR(0) } else { OFQ_DISPATCH_PAIR(1) }
#undef OFQ_DISPATCH_PAIR
    }
    // long K loops: a 4-stage operand ring with one store staging buffer per epilogue warp (5-10 % faster from K ~ 1024);
    // short ones (the K = 384 layers) are epilogue-paced and keep two staging buffers and 3 stages
    const bool deep = (long long)p.kblocks * k2 / splits >= 8;
#define OFQ_DISPATCH(KIND)                                                   \
    switch (bn) {                                                            \
        case 256: return launch_gemm<KIND, 256, 3>(tmA, tmB, tmC, p, st);    \
        case 224: return deep ? launch_gemm<KIND, 224, 4, 1, 1>(tmA, tmB, tmC, p, st)   \
                              : launch_gemm<KIND, 224, 3>(tmA, tmB, tmC, p, st);         \
        case 192: return deep ? launch_gemm<KIND, 192, 4, 1, 1>(tmA, tmB, tmC, p, st)   \
                              : launch_gemm<KIND, 192, 3>(tmA, tmB, tmC, p, st);         \
        case 128: return launch_gemm<KIND, 128, 4>(tmA, tmB, tmC, p, st);    \
        case 64:  return launch_gemm<KIND, 64, 6>(tmA, tmB, tmC, p, st);     \
        default:  return launch_gemm<KIND, 32, 6>(tmA, tmB, tmC, p, st);     \
    }
    if (A->dual_delta > 0) {
        if (kind == OFQ_GEMM_I8 || B->dual_delta != 0 || A->k2_stride == 0) {
            ofq_set_error("ofq_gemm: dual_delta is for a bf16 A operand with an outer-K stride");
            return OFQ_ERR_ARG;
        }
        switch (bn) {
            case 256: return launch_gemm<1, 256, 2, 2, 2>(tmA, tmB, tmC, p, st);
            case 224: return launch_gemm<1, 224, 3, 2, 1>(tmA, tmB, tmC, p, st);
            case 192: return launch_gemm<1, 192, 3, 2, 1>(tmA, tmB, tmC, p, st);
            case 128: return launch_gemm<1, 128, 3, 2, 2>(tmA, tmB, tmC, p, st);
            case 64:  return launch_gemm<1, 64, 3, 2, 2>(tmA, tmB, tmC, p, st);
            default:  return launch_gemm<1, 32, 4, 2, 2>(tmA, tmB, tmC, p, st);
        }
    }
    if (kind == OFQ_GEMM_I8) { OFQ_DISPATCH(0) } else { OFQ_DISPATCH(1) }
#undef OFQ_DISPATCH
}

// ------------------------------------------------------------------------------------------ ofq_gemm_lsq
template <int BN, int STAGES, int MODE, bool RT1>
static int launch_gemm_lsq(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmQ8, const CUtensorMap& tmQ16,
                           const CUtensorMap& tmR16, GemmParams p, const LsqEpi& e, cudaStream_t stream) {
    constexpr size_t smem = LsqSmem<BN, STAGES>::DYN_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_lsq_kernel<BN, STAGES, MODE, RT1>;
    static bool configured = false;
    if (!configured) {
        OFQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int mtiles = (p.M + BM - 1) / BM, ntiles = (p.N + BN - 1) / BN;
    p.fd_ntiles = make_fastdiv(ntiles); p.fd_mtiles = make_fastdiv(mtiles);
    const long long tiles = (long long)mtiles * ntiles;
    OFQ_REQUIRE(tiles <= 0x7fffffff, "ofq_gemm_lsq: too many tiles");
    const int grid = (int)(tiles < ofq_num_sms() ? tiles : ofq_num_sms());
    kern<<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, tmQ8, tmQ16, tmR16, p, e, (int)tiles, mtiles, ntiles);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

// {N, M, 1, 1, 1} map of a quantizer output with a 32-column x 32-row box
static int make_q_map(CUtensorMap* tm, void* ptr, long long ld, int elem_bytes, bool f16, int M, int N) {
    OFQ_REQUIRE(reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && (ld * elem_bytes) % 16 == 0,
                "ofq_gemm_lsq: outputs must be 16-byte aligned with 16-byte-multiple row pitches");
    const cuuint64_t row_b = (cuuint64_t)ld * elem_bytes;
    cuuint64_t dims[5] = {(cuuint64_t)N, (cuuint64_t)M, 1, 1, 1};
    cuuint64_t strides[4] = {row_b, row_b, row_b, row_b};
    cuuint32_t box[5] = {32, 32, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return ofq_encode_tensor_map_sw(tm, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                    : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16),
                                    5, ptr, dims, strides, box, estr,
                                    elem_bytes == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
}

extern "C" int ofq_gemm_lsq(const ofq_operand_t* A, const ofq_operand_t* B, int M, int N, int K, const ofq_vec_t* rs,
                            const ofq_vec_t* cs, const ofq_vec_t* rt, const ofq_vec_t* ct, const ofq_gemm_lsq_t* q, void* stream) {
    OFQ_REQUIRE(A && B && q && q->codes && q->s_eff && q->inv_s, "ofq_gemm_lsq: null argument");
    OFQ_REQUIRE(M > 0 && N > 0 && K > 0 && N % 16 == 0, "ofq_gemm_lsq: extents must be positive, N a multiple of 16");
    OFQ_REQUIRE(q->period > 0 && q->nseg > 0 && q->seg_len > 0 && q->seg_len % 32 == 0 && (long long)q->nseg * q->seg_len >= N,
                "ofq_gemm_lsq: segments must be multiples of 32 columns and cover N");
    OFQ_REQUIRE(!q->rowdot || (q->dot_u && q->workspace), "ofq_gemm_lsq: rowdot needs dot_u and a workspace");
    OFQ_REQUIRE(A->dual_delta == 0 && !A->mn_major && !B->mn_major, "ofq_gemm_lsq: K-major int8 operands only");
    OFQ_CHECK_ARCH();
    GemmParams p;
    p.M = M; p.N = N;
    p.kblocks = (K + KBYTES - 1) / KBYTES;
    p.k2 = 1; p.splits = 1; p.nb1 = p.nb2 = 1;
    p.a_b1 = p.a_b2 = p.b_b1 = p.b_b2 = p.c_b1 = p.c_b2 = 0;
    p.a_k2 = p.b_k2 = 0; p.a_dual_delta = 0;
    p.a_k2mod = p.b_k2mod = 0x7fffffff;
    p.fd_nbatch = make_fastdiv(1); p.fd_nb1 = make_fastdiv(1); p.fd_splits = make_fastdiv(1);
    p.fd_kblocks = make_fastdiv(p.kblocks); p.fd_ak2mod = make_fastdiv(p.a_k2mod); p.fd_bk2mod = make_fastdiv(p.b_k2mod);
    p.rs = make_vec(rs); p.cs = make_vec(cs); p.rt = make_vec(rt); p.ct = make_vec(ct);
    p.has_rank1 = (rt && rt->ptr) || (ct && ct->ptr);
    p.atomic = 0; p.ab_fmt = 0; p.a_mn = p.b_mn = 0; p.amax = nullptr;
    static const int nostore = [] { const char* ev = getenv("OFQ_GEMM_NOSTORE"); return ev ? atoi(ev) : 0; }();
    p.debug_nostore = nostore;      // measurement only: 1 = no TMA stores, 2 = no epilogue math, 8 = no row-dot partial stores
    LsqEpi e;
    e.b4 = q->b4; e.s_eff = q->s_eff; e.inv_s = q->inv_s; e.dot_u = q->rowdot ? q->dot_u : nullptr;
    e.part = q->rowdot ? q->workspace : nullptr;
    e.ld_part = (N + 31) / 32;
    e.nseg = q->nseg; e.seg_len = q->seg_len;
    e.fd_period = make_fastdiv(q->period); e.fd_seglen = make_fastdiv(q->seg_len);
    e.qlo = q->qlo; e.qhi = q->qhi;
    e.has_res = q->res16 != nullptr;
    e.has16 = q->codes16 != nullptr;
    e.f16 = q->fmt16 == OFQ_FMT_F16;
    OFQ_REQUIRE(!e.has_res || e.has16, "ofq_gemm_lsq: the residual is produced together with the 16-bit copy of the codes");
    // tile width: at most one segment boundary per tile
    int bn = q->seg_len >= 192 ? 192 : (q->seg_len >= 128 ? 128 : 64);
    if (N <= 64) bn = 64; else if (N <= 128 && bn > 128) bn = 128;
    CUtensorMap tmA, tmB, tmQ8, tmQ16, tmR16;
    int rc = make_operand_map(&tmA, A, 1, M, K, 1, 1, 1, BM);
    if (rc) return rc;
    rc = make_operand_map(&tmB, B, 1, N, K, 1, 1, 1, bn);
    if (rc) return rc;
    rc = make_q_map(&tmQ8, q->codes, q->ld_codes, 1, false, M, N);
    if (rc) return rc;
    tmQ16 = tmQ8; tmR16 = tmQ8;
    if (e.has16) { rc = make_q_map(&tmQ16, q->codes16, q->ld_codes16, 2, e.f16 != 0, M, N); if (rc) return rc; }
    if (e.has_res) { rc = make_q_map(&tmR16, q->res16, q->ld_res16, 2, true, M, N); if (rc) return rc; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mode = e.has_res ? 2 : (e.has16 ? 1 : 0);
#define OFQ_LSQ_DISPATCH2(BN_, ST_, RT1_)                                                                  \
    (mode == 2 ? launch_gemm_lsq<BN_, ST_, 2, RT1_>(tmA, tmB, tmQ8, tmQ16, tmR16, p, e, st)                \
               : (mode == 1 ? launch_gemm_lsq<BN_, ST_, 1, RT1_>(tmA, tmB, tmQ8, tmQ16, tmR16, p, e, st)   \
                            : launch_gemm_lsq<BN_, ST_, 0, RT1_>(tmA, tmB, tmQ8, tmQ16, tmR16, p, e, st)))
    // RT1: no per-row offset vector (rt = 1): the rank-1 term is ct[n] itself, one multiply less per element
#define OFQ_LSQ_DISPATCH(BN_, ST_) (rt1 ? OFQ_LSQ_DISPATCH2(BN_, ST_, true) : OFQ_LSQ_DISPATCH2(BN_, ST_, false))
    const bool rt1 = !(rt && rt->ptr);
    switch (bn) {
        case 192: rc = OFQ_LSQ_DISPATCH(192, 3); break;
        case 128: rc = OFQ_LSQ_DISPATCH(128, 3); break;
        default:  rc = OFQ_LSQ_DISPATCH(64, 4); break;
    }
#undef OFQ_LSQ_DISPATCH2
#undef OFQ_LSQ_DISPATCH
    if (rc) return rc;
    if (q->rowdot) {
        const long long n = (long long)M * q->nseg;
        rowdot_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e.part, e.ld_part, M, q->nseg, q->seg_len / 32, q->rowdot);
        OFQ_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ ofq_gemm_dx_lsq
template <int BN, int STAGES>
static int launch_gemm_dxlsq(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const CUtensorMap& tmC,
                             GemmParams p, const DxLsqEpi& e, cudaStream_t stream) {
    constexpr size_t smem = DxLsqSmem<BN, STAGES>::DYN_BYTES;
    static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
    auto kern = gemm_dxlsq_kernel<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        OFQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int mtiles = (p.M + BM - 1) / BM, ntiles = (p.N + BN - 1) / BN;
    p.fd_ntiles = make_fastdiv(ntiles); p.fd_mtiles = make_fastdiv(mtiles);
    const long long tiles = (long long)mtiles * ntiles;
    OFQ_REQUIRE(tiles <= 0x7fffffff, "ofq_gemm_dx_lsq: too many tiles");
    const int grid = (int)(tiles < ofq_num_sms() ? tiles : ofq_num_sms());
    kern<<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, tmX, tmC, p, e, (int)tiles, mtiles, ntiles);
    OFQ_CUDA(cudaGetLastError());
    return 0;
}

static int dxlsq_bn(int N) { return N % 192 == 0 ? 192 : (N % 128 == 0 ? 128 : 64); }

extern "C" long long ofq_gemm_dx_lsq_workspace(int M, int N) {
    const int bn = dxlsq_bn(N);
    const long long ntiles = (N + bn - 1) / bn, mtiles = (M + BM - 1) / BM;
    return 2 * ntiles * (long long)M + 4 * mtiles * 3 * (long long)N;
}

extern "C" int ofq_gemm_dx_lsq(int kind, const ofq_operand_t* A, const ofq_operand_t* B, int M, int N, int K, const ofq_vec_t* rs,
                               const ofq_vec_t* cs, const float* x, long long ldx, const float* b4, int qlo, int qhi, float g,
                               float* dx, long long lddx, float* d_s, float* d_b4, float* d_aft, const int8_t* w_codes,
                               long long ld_codes, const float* dy_colsum, const float* colscale, float* dx_absmax, float* workspace,
                               void* stream) {
    OFQ_REQUIRE(kind == OFQ_GEMM_BF16 || kind == OFQ_GEMM_F16, "ofq_gemm_dx_lsq: 16-bit kinds only");
    OFQ_REQUIRE(A && B && rs && rs->ptr && rs->period > 0 && x && dx && workspace && d_b4, "ofq_gemm_dx_lsq: null argument");
    OFQ_REQUIRE(M > 0 && N > 0 && K > 0 && N % 64 == 0, "ofq_gemm_dx_lsq: extents must be positive, N a multiple of 64");
    OFQ_REQUIRE(M % rs->period == 0 || rs->period >= M, "ofq_gemm_dx_lsq: rows must be a multiple of the step-size period");
    OFQ_REQUIRE(A->dual_delta == 0 && A->bstride1 == 0 && B->bstride1 == 0, "ofq_gemm_dx_lsq: plain (unbatched, single-plane) operands only");
    OFQ_CHECK_ARCH();
    const int kelem = KBYTES / 2;
    GemmParams p;
    p.M = M; p.N = N;
    p.kblocks = (K + kelem - 1) / kelem;
    p.k2 = 1; p.splits = 1; p.nb1 = p.nb2 = 1;
    p.a_b1 = p.a_b2 = p.b_b1 = p.b_b2 = p.c_b1 = p.c_b2 = 0;
    p.a_k2 = p.b_k2 = 0; p.a_dual_delta = 0;
    p.a_k2mod = p.b_k2mod = 0x7fffffff;
    p.fd_nbatch = make_fastdiv(1); p.fd_nb1 = make_fastdiv(1); p.fd_splits = make_fastdiv(1);
    p.fd_kblocks = make_fastdiv(p.kblocks); p.fd_ak2mod = make_fastdiv(p.a_k2mod); p.fd_bk2mod = make_fastdiv(p.b_k2mod);
    p.rs = make_vec(rs); p.cs = make_vec(cs); p.rt = make_vec(nullptr); p.ct = make_vec(nullptr);
    p.has_rank1 = 0; p.atomic = 0;
    p.ab_fmt = kind == OFQ_GEMM_F16 ? 0 : 1;
    p.a_mn = A->mn_major != 0; p.b_mn = B->mn_major != 0;
    p.amax = nullptr; p.debug_nostore = 0;
    const int bn = dxlsq_bn(N);
    OFQ_REQUIRE(!p.b_mn || bn % 64 == 0, "ofq_gemm_dx_lsq: tile width");
    const long long ntiles = (N + bn - 1) / bn, mtiles = (M + BM - 1) / BM;
    DxLsqEpi e;
    e.amax = reinterpret_cast<unsigned int*>(dx_absmax);
    e.b4 = b4; e.rowpart = workspace; e.colpart = workspace + 2 * ntiles * (long long)M;
    e.qlo = (float)qlo; e.qhi = (float)qhi;
    CUtensorMap tmA, tmB, tmX, tmC;
    int rc = make_operand_map(&tmA, A, 2, M, K, 1, 1, 1, BM, kind == OFQ_GEMM_F16);
    if (rc) return rc;
    rc = make_operand_map(&tmB, B, 2, N, K, 1, 1, 1, bn, kind == OFQ_GEMM_F16);
    if (rc) return rc;
    ofq_gemm_out_t ox = {const_cast<float*>(x), ldx, 0, 0, 0}, oc = {dx, lddx, 0, 0, 0};
    rc = make_out_map(&tmX, &ox, M, N, 1, 1);
    if (rc) return rc;
    rc = make_out_map(&tmC, &oc, M, N, 1, 1);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (bn) {
        case 192: rc = launch_gemm_dxlsq<192, 3>(tmA, tmB, tmX, tmC, p, e, st); break;
        case 128: rc = launch_gemm_dxlsq<128, 3>(tmA, tmB, tmX, tmC, p, e, st); break;
        default:  rc = launch_gemm_dxlsq<64, 4>(tmA, tmB, tmX, tmC, p, e, st); break;
    }
    if (rc) return rc;
    // d_aft[k] = sum_m dX_hat[m][k] = sum_n colsum(dY)[n] * colscale[n] * W_codes[n][k]  (exact identity; no pass over dX_hat)
    if (d_aft) {
        OFQ_REQUIRE(w_codes && dy_colsum, "ofq_gemm_dx_lsq: d_aft needs the int8 weight codes and colsum(dY)");
        codes_vecmat_parts_kernel<<<(unsigned)(4 * mtiles), 128, 0, st>>>(w_codes, K, N, ld_codes, dy_colsum, colscale, e.colpart, (int)(4 * mtiles));
        OFQ_CUDA(cudaGetLastError());
    }
    return ofq_lsq_bwd_finalize_parts(e.colpart, 4 * mtiles, e.rowpart, 2 * ntiles * (long long)M, M, N, rs->period, g, d_s, d_b4, d_aft, stream);
}
