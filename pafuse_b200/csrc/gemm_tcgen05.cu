// Token-wise linear layers of the STE/TTE blocks (mixste.py:65,80,38-41) on the
// sm_100a tensor cores:  D[M,N] = A[M,K] * W[N,K]^T + bias, fp32 accumulation in TMEM.
//
// Precision: "f16x3".  Every fp32 operand v is carried as the pair
// (hi, lo) = (fp16(v), fp16(v - hi)) and the product is evaluated as
//     A_lo*W_hi + A_hi*W_lo + A_hi*W_hi          (3 tcgen05.mma per K-slice)
// (~22 significant bits per operand; plain TF32/BF16 do not meet the path's
// tolerance).  Weights are stored pre-scaled by 2^8 (common.cuh).
//
// Structure: persistent, warp-specialised, one CTA per SM; with CG == 2 the two CTAs
// of a cluster form a tcgen05 CTA pair that owns a 256 x BN output tile: each CTA
// TMA-loads its own 128 rows of A and HALF of the W tile (BN/2 rows), the even CTA
// issues cta_group::2 MMAs that read both shared memories and write both tensor
// memories.  Per byte fetched from L2 the pair does twice the math of a lone CTA --
// the 1-CTA version of this kernel was L2->SMEM bound at 42 % tensor utilisation
// (profiles/r1a_gemm_1cta_bf16x3_ncu_full.txt).
//   warp 0    TMA producer: 4 tiled loads per stage (A_hi, A_lo, W_hi, W_lo; 128B swizzle)
//   warp 1    MMA issuer (even CTA only): 3 x (BK/16) tcgen05.mma per stage, commits to mbarriers
//   warp 2    TMEM allocator (512 columns = two BN<=256 accumulator buffers)
//   warps 4-19 epilogue, FOUR warps per TMEM lane quarter, each taking every fourth 16-column chunk: tcgen05.ld
//             32 lanes x 16 columns, bias / GELU+split / residual + LayerNorm, swizzled st.shared into a per-warp
//             2 KB staging box, TMA store of the box (cp.reduce.async.bulk .add for the residual update x += y,
//             performed at the L2).  Registers: setmaxnreg moves the control warpgroup down to REGS_CTRL and the four
//             epilogue warpgroups up to REGS_EPI.
//             History: v1 stored straight from registers, one row per lane: 32 different 128-byte lines per store
//             instruction made the epilogue the critical path (profiles/r1a_*); v2 had 4 epilogue warps and the
//             GELU / head-plane epilogues (40 instructions per element at 0.4 IPC) were still slower than the main
//             loop (profiles/r1b_gemm_attention_ncu_full.txt); v3 (round 1) had 8 warps on 32-column chunks: every
//             epilogue is a serial chain tcgen05.ld -> math -> st.shared -> TMA store per chunk, and with two warps
//             per scheduler those chains ran at IPC ~0.2 and paced the face / hands launches (fc1 and qkv issue
//             bound, proj / fc2 at 5 TB/s, profiles/r1n_epilogue_analysis.txt).  Sixteen warps give every scheduler
//             four independent chains; 16-column chunks keep the staging at 32 KB and the per-thread state small
//             enough for 640 threads.
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty pair (MMA <-> epilogue),
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include "kernels.cuh"

#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

namespace pafuse {

namespace {

constexpr int BM = 128;              // rows per CTA = TMEM lanes
constexpr int BKW = 64;              // K extent of a W box = one 128-byte swizzle span of fp16
constexpr int UK = 16;               // K per tcgen05.mma (16-bit operands)
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;
// static shared memory (barriers; + the 2 KB statistics exchange buffer of EPI_RESID_LN) rounded up to the 1024-byte
// alignment of the dynamic window that follows it
constexpr int SMEM_SLACK = 1024;
constexpr int SMEM_SLACK_LN = 3072;
constexpr int EPI_WARP0 = 4;
#ifndef PAFUSE_EPI_WARPS
#define PAFUSE_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = PAFUSE_EPI_WARPS;         // 16 (default) or 8 (A/B builds: -DPAFUSE_EPI_WARPS=8)
constexpr int EPI_SUBS = EPI_WARPS / 4;             // epilogue warps per TMEM lane quarter
constexpr int NUM_THREADS = (EPI_WARP0 + EPI_WARPS) * 32;   // 640
constexpr int CW = 16;                              // accumulator columns per epilogue chunk
constexpr int TMEM_COLS = 512;
constexpr int STG_WARP_BYTES = 2048;                // per epilogue warp: one staging box of 32 rows x 64 B (fp32) or hi + lo boxes of 32 rows x 32 B
constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
// register budget (setmaxnreg, per thread): 4 control warps x 32 x 64 + 16 epilogue warps x 32 x 104 = 61440 <= 65536
#if PAFUSE_EPI_WARPS == 16
#define PAFUSE_REGS_CTRL "64"
#define PAFUSE_REGS_EPI "104"
#else                                               // 384 threads x 168: 4 x 32 x 64 + 8 x 32 x 216 = 63488 <= 64512
#define PAFUSE_REGS_CTRL "64"
#define PAFUSE_REGS_EPI "216"
#endif

struct KernelParams {
    long long M;
    int N, K;
    int block_n;                     // output tile width (UMMA N)
    int m_tiles, n_tiles;            // tiles of (BM*CG) x block_n
    int stages;
    int stage_bytes;                 // per CTA
    int a_bk;                        // K extent of an A stage: 64 (128-byte swizzle) or 32 (64-byte swizzle)
    int w_res_bytes;                 // weight-stationary mode: bytes of the resident W slice per CTA
    int slots;                       // weight-stationary mode: groups of n_tiles CTA pairs
    float out_scale;                 // undoes WEIGHT_SCALE
    const float* bias;
    int hds;                         // EPI_PLANES: stored head width (columns per plane, a multiple of 16)
    GemmLnFuse ln;                   // EPI_RESID_LN
#ifdef PAFUSE_ABLATE
    int ablate;                      // energy ablation builds only (tools/energy_ablation.py): 1 = epilogue warps skip their work, 2 = no MMAs
#endif
};

// Tile walk of one CTA group.  Streaming mode: tiles (m, n) in m-major order, strided over the groups, so
// consecutive groups share an A tile through the L2.  Weight-stationary mode (WRES): the group keeps one n tile
// for the whole launch (its W slice stays in shared memory) and walks m tiles; the n_tiles groups of a "slot"
// walk the same m tiles at the same time, so every A tile still comes from DRAM once.
template <bool WRES>
struct TileWalk {
    int first, step, count, n_fixed, n_tiles;
    __device__ TileWalk(const KernelParams& p, int group, int num_groups) {
        n_tiles = p.n_tiles;
        if (WRES) {
            const int slot = group / p.n_tiles;
            n_fixed = group % p.n_tiles;
            first = slot;
            step = p.slots;
            count = slot < p.slots ? (p.m_tiles - slot + p.slots - 1) / p.slots : 0;
        } else {
            const int total = p.m_tiles * p.n_tiles;
            n_fixed = 0;
            first = group;
            step = num_groups;
            count = group < total ? (total - group + num_groups - 1) / num_groups : 0;
        }
    }
    __device__ void at(int i, int& m_tile, int& n_tile) const {
        const int t = first + i * step;
        if (WRES) {
            m_tile = t;
            n_tile = n_fixed;
        } else {
            m_tile = t / n_tiles;
            n_tile = t % n_tiles;
        }
    }
};

// 32 lanes x 16 consecutive 32-bit columns <-> 16 registers per thread (thread i <-> lane base + i)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait_all() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#ifdef PAFUSE_GEMM_TRACE
__device__ long long g_gemm_tr_bulk;   // measurement builds only: cycles lane 0 of the first epilogue warp of CTA 0 waits for its staging box
#define GEMM_BULK_WAIT_READ()                                                                  \
    {                                                                                          \
        const long long t_bw = clock64();                                                      \
        bulk_wait_group_read<0>();                                                             \
        if (blockIdx.x == 0 && (threadIdx.x >> 5) == 4) g_gemm_tr_bulk += clock64() - t_bw;    \
    }
__device__ long long g_gemm_tr_ln[16];  // LayerNorm epilogue phases of the same warp: pass 1, merge 1, pass 2, merge 2, pass 3
#define LN_TR(i)                                                                               \
    if (blockIdx.x == 0 && threadIdx.x == 4 * 32) {                                            \
        const long long t_now = clock64();                                                     \
        g_gemm_tr_ln[i] += t_now - ln_t_last;                                                  \
        ln_t_last = t_now;                                                                     \
    }
#else
#define GEMM_BULK_WAIT_READ() bulk_wait_group_read<0>()
#define LN_TR(i)
#endif
// the EPI_SUBS warps that share a TMEM lane quarter (128 threads), barrier id 1 + quarter
__device__ __forceinline__ void quarter_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(EPI_SUBS * 32) : "memory"); }
// staging-box offsets of 16-byte piece i of row `row`: fp32 boxes are 32 rows x 64 B with the 64-byte swizzle
// (piece i at i ^ ((row >> 1) & 3)), fp16 boxes 32 rows x 32 B with the 32-byte swizzle (piece i at i ^ ((row >> 2) & 1));
// both patterns are bank-conflict free for a quarter-warp of consecutive rows
__device__ __forceinline__ uint32_t box_off_f32(int row, int i) { return (uint32_t)(row * 64 + ((i ^ ((row >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint32_t box_off_f16(int row, int i) { return (uint32_t)(row * 32 + ((i ^ ((row >> 2) & 1)) << 4)); }
// accumulator registers 2j, 2j+1 as a packed pair
__device__ __forceinline__ float2 pair_of(const uint32_t (&r)[16], int j) {
    return make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
}
// 16 values (8 pairs) -> fp16 hi / lo boxes (hi at +0, lo at +1024) of this warp's staging area
__device__ __forceinline__ void stage_split16(uint32_t box_s, int lane, const float2 (&v)[8]) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_pair_sat2(v[i], hi[i], lo[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t off = box_off_f16(lane, i);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box_s + off), "r"(hi[4 * i]), "r"(hi[4 * i + 1]),
                     "r"(hi[4 * i + 2]), "r"(hi[4 * i + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box_s + 1024 + off), "r"(lo[4 * i]), "r"(lo[4 * i + 1]),
                     "r"(lo[4 * i + 2]), "r"(lo[4 * i + 3]) : "memory");
    }
}

// Epilogue of the residual GEMMs (proj, fc2) of the parts whose rows fit one tile (N == BN <= 256), fused with the
// LayerNorms that follow them (mixste.py:114-115 norm2; :243,257,269,273 the shared norm that closes the block,
// :250 Temporal_pos_embed, and norm1 of the next block):
//     v  = x + acc * scale + bias                                   (x: fp32 residual stream, read here)
//     not chained:  x <- v ;                          a = LN(v; g1, b1)   -> fp16 hi/lo
//     chained:      x <- y = LN(v; g0, b0) [+ add_f[f]] ;  a = LN(y; g1, b1)   -> fp16 hi/lo
// It replaces the L2-side reduction of EPI_RESID plus one ln_chain_kernel launch (8-12 B per element of DRAM
// traffic).  A thread owns one row (= TMEM lane); the EPI_SUBS warps of a lane quarter take every EPI_SUBS-th 16-column
// chunk and merge their partial row statistics through shared memory (one exchange per norm).  The row values stay in
// the accumulator's tensor memory between the passes (tcgen05.st / tcgen05.ld); statistics are gathered in the pass that
// produces the values (shifted sums, see merge_stats).  x is fetched with coalesced 16-byte loads one chunk ahead
// (registers), transposed through the warp's staging box, which then carries the output row by row to the TMA store.
// coalesced fetch of a 32 x 16 box of x: lane -> 16 bytes at column (lane & 3) * 4 of rows (lane >> 2) + 8 i
// x of one 16-column chunk, loaded coalesced (four lanes per row) and transposed through the staging box by the caller.
// (Thread = row loads of the lane's own 64 bytes, without the transposition, measured SLOWER: the 32-line loads congest
// the LSU and every other pass of the epilogue pays for it -- GEMM time per step 805 -> 829 ms, profiles/r2ag_*.)
__device__ __forceinline__ void ln_fetch_x(const KernelParams& p, float4 (&xr)[4], int row0, int c0, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = (long long)row0 + (lane >> 2) + 8 * i;
        xr[i] = r < p.M ? __ldg(reinterpret_cast<const float4*>(p.ln.x + (size_t)r * p.N + c0 + (lane & 3) * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// 16 consecutive floats of a per-column parameter vector, the same for every lane (broadcast loads).  Issued as one
// batch ahead of their use: loaded one by one next to the FMAs they feed, every load's latency was exposed
// (profiles/r1j_*: a quarter of the epilogue time).
__device__ __forceinline__ void load_vec16(float4 (&v)[4], const float* src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldg(s4 + i);
}

// (A bulk L2 prefetch of the next tile's residual rows -- they are one contiguous range of x -- was tried against the
// x-fetch stalls of this epilogue and made proj / fc2 5-14 % slower, with 25 % more DRAM reads: profiles/r1n_*.)
template <int CG>
__device__ __forceinline__ void epilogue_resid_ln(const KernelParams& p, const CUtensorMap& tm_x, const CUtensorMap& tm_hi,
                                                  const CUtensorMap& tm_lo, uint8_t* box, float* xch, uint32_t t_base,
                                                  int row0, int q, int sub, int lane, int BN, float oscale,
                                                  uint64_t* tmem_empty, float4 (&xr)[4]) {
    const GemmLnFuse& f = p.ln;
    const int N = p.N;
    const int nck = BN / CW;
    const int mine = (nck - sub + EPI_SUBS - 1) / EPI_SUBS;           // chunks sub, sub + EPI_SUBS, ... (may be 0)
    const bool chained = f.g0 != nullptr;
    const float invN = 1.0f / (float)N;
    float* my_x = xch + (q * 32 + lane) * EPI_SUBS;
    const int bar_id = 1 + q;
    const uint32_t box_s = smem_u32(box);

    // Row statistics in ONE pass over the values (a second tensor-memory pass per norm cost 40 % of the epilogue):
    // each warp accumulates S = sum(v - K), Q = sum((v - K)^2) around K = its first value of the row, which keeps
    // Q - S^2/n free of cancellation; the warps' (mean, M2) are merged with the pairwise update of Chan et al., by
    // every warp in the same order, so all of them hold bit-identical statistics.
    // The exchange buffer holds ONE float per (row, warp) -- 2 KB, which is what the operand ring leaves at C = 256 --
    // so means and M2s travel in two rounds.
    auto exchange = [&](float mine_v, float (&all)[EPI_SUBS]) {
        my_x[sub] = mine_v;
        quarter_bar_sync(bar_id);
#pragma unroll
        for (int s = 0; s < EPI_SUBS; ++s) all[s] = my_x[s];
        quarter_bar_sync(bar_id);                                     // the slots may be rewritten after this point
    };
    auto merge_stats = [&](float K, float S, float Q, float eps, float& mean, float& rstd) {
        const float na = (float)(mine * CW);
        float means[EPI_SUBS], m2s[EPI_SUBS];
        exchange(mine > 0 ? K + S / na : 0.f, means);
        exchange(mine > 0 ? Q - S * S / na : 0.f, m2s);
        float n_acc = 0.f, mean_acc = 0.f, m2_acc = 0.f;
#pragma unroll
        for (int s = 0; s < EPI_SUBS; ++s) {
            const int cnt = (nck - s + EPI_SUBS - 1) / EPI_SUBS;
            if (cnt > 0) {                                            // warp-uniform
                const float n_s = (float)(cnt * CW), n_new = n_acc + n_s;
                const float delta = means[s] - mean_acc;
                mean_acc += delta * (n_s / n_new);
                m2_acc += m2s[s] + delta * delta * (n_acc * n_s / n_new);
                n_acc = n_new;
            }
        }
        mean = mean_acc;
        rstd = 1.0f / sqrtf(m2_acc * invN + eps);
    };

#ifdef PAFUSE_GEMM_TRACE
    long long ln_t_last = clock64();
#endif
    // ---- pass 1: v = x + acc * scale + bias -> tensor memory (and, when not chained, -> x)
    // (all row arithmetic on packed pairs: FFMA2 / FADD2, see common.cuh)
    const float2 osc2 = splat2(oscale);
    float K = 0.f;
    float2 S2 = splat2(0.f), Q2 = splat2(0.f);
    for (int ci = 0; ci < mine; ++ci) {
        const int c0 = (sub + EPI_SUBS * ci) * CW;
        uint32_t r[16];
        float4 bv[4];
        tmem_ld_32x16(t_base + (uint32_t)c0, r);
        load_vec16(bv, p.bias + c0);
        if (lane == 0) GEMM_BULK_WAIT_READ();                     // the store that last used the box has read it
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i)                                   // transpose x through the box
            sts128(box_s + box_off_f32((lane >> 2) + 8 * i, lane & 3), xr[i]);
        __syncwarp();
        if (ci + 1 < mine) ln_fetch_x(p, xr, row0, c0 + EPI_SUBS * CW, lane);   // next chunk's x, in flight during the math
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t slot = box_s + box_off_f32(lane, i);
            const float4 xv = lds128(slot);
            const float2 va = fadd2(ffma2(pair_of(r, 2 * i), osc2, make_float2(bv[i].x, bv[i].y)), make_float2(xv.x, xv.y));
            const float2 vb = fadd2(ffma2(pair_of(r, 2 * i + 1), osc2, make_float2(bv[i].z, bv[i].w)), make_float2(xv.z, xv.w));
            if (ci == 0 && i == 0) K = va.x;
            const float2 nK = splat2(-K);
            const float2 da = fadd2(va, nK), db = fadd2(vb, nK);
            S2 = fadd2(S2, fadd2(da, db));
            Q2 = ffma2(da, da, ffma2(db, db, Q2));
            r[4 * i + 0] = __float_as_uint(va.x);
            r[4 * i + 1] = __float_as_uint(va.y);
            r[4 * i + 2] = __float_as_uint(vb.x);
            r[4 * i + 3] = __float_as_uint(vb.y);
            if (!chained) sts128(slot, make_float4(va.x, va.y, vb.x, vb.y));
        }
        tmem_st_32x16(t_base + (uint32_t)c0, r);
        if (!chained) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_x, box, c0, row0);
                bulk_commit_group();
            }
        }
    }
    tmem_st_wait_all();
    LN_TR(0)
    float mean, rstd;
    merge_stats(K, S2.x + S2.y, Q2.x + Q2.y, chained ? f.eps0 : f.eps1, mean, rstd);
    LN_TR(1)

    if (chained) {
        // ---- y = LN(v; g0, b0) [+ add_f[f]] -> x and tensor memory
        const float2 rstd2 = splat2(rstd), mr2 = splat2(-mean * rstd);
        const long long row = (long long)row0 + lane;
        const float* addr = f.add_f ? f.add_f + (size_t)((row / f.J) % f.F) * N : nullptr;
        K = 0.f; S2 = splat2(0.f); Q2 = splat2(0.f);
        for (int ci = 0; ci < mine; ++ci) {
            const int c0 = (sub + EPI_SUBS * ci) * CW;
            uint32_t r[16];
            float4 gv[4], bv[4];
            tmem_ld_32x16(t_base + (uint32_t)c0, r);
            load_vec16(gv, f.g0 + c0);
            load_vec16(bv, f.b0 + c0);
            tmem_ld_wait();
            float2 y[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[2 * i] = ffma2(ffma2(pair_of(r, 2 * i), rstd2, mr2), make_float2(gv[i].x, gv[i].y), make_float2(bv[i].x, bv[i].y));
                y[2 * i + 1] = ffma2(ffma2(pair_of(r, 2 * i + 1), rstd2, mr2), make_float2(gv[i].z, gv[i].w), make_float2(bv[i].z, bv[i].w));
            }
            if (addr) {                                               // Temporal_pos_embed (after STE block 0 only)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(addr + c0) + i);
                    y[2 * i] = fadd2(y[2 * i], make_float2(a.x, a.y));
                    y[2 * i + 1] = fadd2(y[2 * i + 1], make_float2(a.z, a.w));
                }
            }
            if (lane == 0) GEMM_BULK_WAIT_READ();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (ci == 0 && i == 0) K = y[0].x;
                const float2 nK = splat2(-K);
                const float2 da = fadd2(y[2 * i], nK), db = fadd2(y[2 * i + 1], nK);
                S2 = fadd2(S2, fadd2(da, db));
                Q2 = ffma2(da, da, ffma2(db, db, Q2));
                sts128(box_s + box_off_f32(lane, i), make_float4(y[2 * i].x, y[2 * i].y, y[2 * i + 1].x, y[2 * i + 1].y));
                r[4 * i + 0] = __float_as_uint(y[2 * i].x);
                r[4 * i + 1] = __float_as_uint(y[2 * i].y);
                r[4 * i + 2] = __float_as_uint(y[2 * i + 1].x);
                r[4 * i + 3] = __float_as_uint(y[2 * i + 1].y);
            }
            tmem_st_32x16(t_base + (uint32_t)c0, r);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_x, box, c0, row0);
                bulk_commit_group();
            }
        }
        tmem_st_wait_all();
        LN_TR(2)
        merge_stats(K, S2.x + S2.y, Q2.x + Q2.y, f.eps1, mean, rstd);
        LN_TR(3)
    }

    // ---- a = LN(row values in tensor memory; g1, b1) -> fp16 hi/lo boxes -> TMA stores
    {
        const float2 rstd2 = splat2(rstd), mr2 = splat2(-mean * rstd);
        for (int ci = 0; ci < mine; ++ci) {
            const int c0 = (sub + EPI_SUBS * ci) * CW;
            uint32_t r[16];
            float4 gv[4], bv[4];
            tmem_ld_32x16(t_base + (uint32_t)c0, r);
            load_vec16(gv, f.g1 + c0);
            load_vec16(bv, f.b1 + c0);
            tmem_ld_wait();
            float2 v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                v[2 * i] = ffma2(ffma2(pair_of(r, 2 * i), rstd2, mr2), make_float2(gv[i].x, gv[i].y), make_float2(bv[i].x, bv[i].y));
                v[2 * i + 1] = ffma2(ffma2(pair_of(r, 2 * i + 1), rstd2, mr2), make_float2(gv[i].z, gv[i].w), make_float2(bv[i].z, bv[i].w));
            }
            if (lane == 0) GEMM_BULK_WAIT_READ();                 // the store that last used the box has read it
            __syncwarp();
            stage_split16(box_s, lane, v);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_hi, box, c0, row0);
                tma_store_2d(&tm_lo, box + 1024, c0, row0);
                bulk_commit_group();
            }
        }
    }

    LN_TR(4)
    // the accumulator buffer goes back to the MMA issuer
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (CG == 1) mbar_arrive(tmem_empty);
        else mbar_arrive_cluster(tmem_empty, 0);
    }
}

template <int EPI, int CG, bool WRES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                  const __grid_constant__ CUtensorMap tm_out0, const __grid_constant__ CUtensorMap tm_out1,
                  const __grid_constant__ CUtensorMap tm_out2, const KernelParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];           // the operand tiles need 1024-byte boundaries (128B swizzle atoms)
    __shared__ float ln_xch[EPI == EPI_RESID_LN ? 128 * EPI_SUBS : 1];   // partial row statistics of the warps that share a row
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t w_full_bar;
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = smem_raw;
    uint8_t* smem_w = smem;                                   // WRES: [K/64][hi, lo][WN rows x 128 B]
    uint8_t* smem_st = smem + (WRES ? p.w_res_bytes : 0);     // operand stage ring
    uint8_t* smem_box = smem_st + (size_t)p.stages * p.stage_bytes;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler: role code uses the uniform datapath
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int BN = p.block_n;
    const int WN = BN / CG;                                   // W rows this CTA loads
    const int w_box_bytes = WN * BKW * 2;
    const int a_bk = WRES ? p.a_bk : BKW;
    const int a_tile_bytes = BM * a_bk * 2;
    const int num_kb = (p.K + a_bk - 1) / a_bk;
    const TileWalk<WRES> walk(p, (int)blockIdx.x / CG, (int)gridDim.x / CG);

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if (smem_u32(smem_raw) & 1023u) {                             // the launch code sizes the ring for an aligned window
            printf("pafuse: dynamic shared memory is not 1024-byte aligned\n");
            asm volatile("trap;");
        }
        prefetch_tensormap(&tm_a_hi);
        prefetch_tensormap(&tm_a_lo);
        prefetch_tensormap(&tm_w_hi);
        prefetch_tensormap(&tm_w_lo);
        prefetch_tensormap(&tm_out0);
        if (EPI == EPI_GELU_SPLIT || EPI == EPI_PLANES || EPI == EPI_RESID_LN) prefetch_tensormap(&tm_out1);
        if (EPI == EPI_RESID_LN) prefetch_tensormap(&tm_out2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);                       // the leader's arrive.expect_tx
            mbar_init(&empty_bar[s], 1);                      // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);                  // one tcgen05.commit
            mbar_init(&tmem_empty_bar[a], EPI_WARPS * CG);    // one elected lane per epilogue warp of every CTA of the group
        }
        mbar_init(&w_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<CG>(&tmem_base_slot, TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();    // peers' barriers are initialised past this point
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    // everything above ran while the preceding kernel was still draining; the weight-stationary producer also
    // fetches its W slice (weights are not written by any kernel of the chain) before it joins the wait
    if (!(WRES && warp == 0)) pdl_wait();

    if (warp < EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " PAFUSE_REGS_CTRL ";");   // the control warpgroup donates registers ...
    if (warp == 0) {
        // ===================== TMA producer (every CTA loads its own operands) =====================
        if (WRES) {
            if (lane == 0 && walk.count > 0) {
                // the whole K extent of this group's W slice, once
                const int n0 = walk.n_fixed * BN + (int)cta_rank * WN;
                const int nkw = (p.K + BKW - 1) / BKW;
                if (CG == 1) mbar_arrive_expect_tx(&w_full_bar, (uint32_t)p.w_res_bytes);
                else if (leader) mbar_arrive_expect_tx(&w_full_bar, (uint32_t)(2 * p.w_res_bytes));
                for (int kw = 0; kw < nkw; ++kw) {
                    uint8_t* dst = smem_w + (size_t)kw * 2 * w_box_bytes;
                    if (CG == 1) {
                        tma_load_2d(dst, &tm_w_hi, &w_full_bar, kw * BKW, n0);
                        tma_load_2d(dst + w_box_bytes, &tm_w_lo, &w_full_bar, kw * BKW, n0);
                    } else {
                        tma_load_2d_pair(dst, &tm_w_hi, &w_full_bar, kw * BKW, n0);
                        tma_load_2d_pair(dst + w_box_bytes, &tm_w_lo, &w_full_bar, kw * BKW, n0);
                    }
                }
            }
            pdl_wait();
        }
        if (lane == 0 && walk.count > 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < walk.count; ++i) {
                int m_tile, n_tile;
                walk.at(i, m_tile, n_tile);
                const int m0 = (m_tile * CG + (int)cta_rank) * BM;
                const int n0 = n_tile * BN + (int)cta_rank * WN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);  // MMAs that read this stage (in both CTAs) have retired
                    uint8_t* st = smem_st + (size_t)stage * p.stage_bytes;
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.stage_bytes);
                        tma_load_2d(st, &tm_a_hi, &full_bar[stage], kb * a_bk, m0);
                        tma_load_2d(st + a_tile_bytes, &tm_a_lo, &full_bar[stage], kb * a_bk, m0);
                        if (!WRES) {
                            tma_load_2d(st + 2 * a_tile_bytes, &tm_w_hi, &full_bar[stage], kb * BKW, n0);
                            tma_load_2d(st + 2 * a_tile_bytes + w_box_bytes, &tm_w_lo, &full_bar[stage], kb * BKW, n0);
                        }
                    } else {
                        // both CTAs' bytes are counted on the leader's barrier
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(2 * p.stage_bytes));
                        tma_load_2d_pair(st, &tm_a_hi, &full_bar[stage], kb * a_bk, m0);
                        tma_load_2d_pair(st + a_tile_bytes, &tm_a_lo, &full_bar[stage], kb * a_bk, m0);
                        if (!WRES) {
                            tma_load_2d_pair(st + 2 * a_tile_bytes, &tm_w_hi, &full_bar[stage], kb * BKW, n0);
                            tma_load_2d_pair(st + 2 * a_tile_bytes + w_box_bytes, &tm_w_lo, &full_bar[stage], kb * BKW, n0);
                        }
                    }
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the group's even CTA) =====================
        if (lane == 0 && leader && walk.count > 0) {
            const uint32_t idesc = make_idesc_f16((uint32_t)(BM * CG), (uint32_t)BN);
            const uint32_t a_layout = a_bk == 64 ? 2u : 4u;           // 128-byte / 64-byte swizzle
            const uint32_t a_sbo = 8u * (uint32_t)a_bk * 2u;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            if (WRES) {
                mbar_wait(&w_full_bar, 0);                            // the resident W slice landed (in both CTAs)
                tcgen05_fence_after();
            }
            for (int i = 0; i < walk.count; ++i) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);       // every epilogue warp of the group drained this buffer
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);               // operands landed (in both CTAs)
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem_st + (size_t)stage * p.stage_bytes);
                    const uint32_t a_hi = sa, a_lo = sa + a_tile_bytes;
                    uint32_t w_hi, w_lo;
                    if (WRES) {
                        const int k0 = kb * a_bk;                     // a_bk divides 64: the stage lies inside one W box
                        w_hi = smem_u32(smem_w) + (uint32_t)(k0 / BKW) * 2u * (uint32_t)w_box_bytes + (uint32_t)(k0 % BKW) * 2u;
                        w_lo = w_hi + w_box_bytes;
                    } else {
                        w_hi = sa + 2 * a_tile_bytes;
                        w_lo = w_hi + w_box_bytes;
                    }
                    int k_left = p.K - kb * a_bk;
                    int nk = k_left >= a_bk ? a_bk / UK : (k_left + UK - 1) / UK;   // K tail: TMA zero-fills, skip dead slices
#ifdef PAFUSE_ABLATE
                    if (p.ablate & 2) nk = 0;                         // operands still travel, barriers still cycle
#endif
                    for (int k = 0; k < nk; ++k) {
                        const uint32_t koff = (uint32_t)k * UK * 2;   // bytes along K inside the swizzle span
                        const uint64_t dah = make_smem_desc(a_hi + koff, a_sbo, a_layout);
                        const uint64_t dal = make_smem_desc(a_lo + koff, a_sbo, a_layout);
                        const uint64_t dwh = make_smem_desc_sw128(w_hi + koff);
                        const uint64_t dwl = make_smem_desc_sw128(w_lo + koff);
                        umma_f16_ss<CG>(d_tmem, dal, dwh, idesc, (kb | k) != 0 ? 1u : 0u);   // small terms first
                        umma_f16_ss<CG>(d_tmem, dah, dwl, idesc, 1u);
                        umma_f16_ss<CG>(d_tmem, dah, dwh, idesc, 1u);
                    }
                    umma_commit<CG>(&empty_bar[stage]);               // frees the stage (in both CTAs) when the MMAs retire
                    if (kb == num_kb - 1) umma_commit<CG>(&tmem_full_bar[acc]);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " PAFUSE_REGS_EPI ";");   // ... to the four epilogue warpgroups
        // ===================== epilogue (every CTA drains its own 128 accumulator rows) =====================
        const int q = warp & 3;                                       // TMEM lane quarter == warp % 4
        const int sub = (warp - EPI_WARP0) >> 2;                      // takes the 16-column chunks sub, sub + EPI_SUBS, ...
        const int nck = BN / CW;
        int acc = 0;
        uint32_t acc_phase = 0;
        const float oscale = p.out_scale;
        uint8_t* box = smem_box + (warp - EPI_WARP0) * STG_WARP_BYTES;   // 1024-byte aligned
        const uint32_t box_s = smem_u32(box);
        auto release = [&](uint64_t* bar) {                           // this warp's share of the accumulator is read
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {                                          // EPI_WARPS*CG arrivals release the buffer to the issuer
                if (CG == 1) mbar_arrive(bar);
                else mbar_arrive_cluster(bar, 0);
            }
        };
#ifdef PAFUSE_GEMM_TRACE
        // measurement builds only: cycles this epilogue warp waits for an accumulator vs. works on it, per tile
        long long tr_wait = 0, tr_work = 0, tr_last = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 4 * 32) {
            g_gemm_tr_bulk = 0;
            for (int i = 0; i < 16; ++i) g_gemm_tr_ln[i] = 0;
        }
#define GEMM_TR(acc) { const long long t_now = clock64(); acc += t_now - tr_last; tr_last = t_now; }
#else
#define GEMM_TR(acc)
#endif
        for (int it = 0; it < walk.count; ++it) {
            int m_tile, n_tile;
            walk.at(it, m_tile, n_tile);
            const int row0 = (m_tile * CG + (int)cta_rank) * BM + q * 32;   // first row of this warp's box
            float4 xr[4];                                             // EPI_RESID_LN: x of the next chunk
#ifdef PAFUSE_ABLATE
            if (p.ablate & 1) {                                       // accumulator handed straight back: no loads, math, stores
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tcgen05_fence_after();
                release(&tmem_empty_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
                continue;
            }
#endif
            if (EPI == EPI_RESID_LN) {
                if (sub < nck) ln_fetch_x(p, xr, row0, sub * CW, lane);   // in flight while the main loop finishes
            }
            GEMM_TR(tr_work)
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tcgen05_fence_after();
            GEMM_TR(tr_wait)
            const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            if (EPI == EPI_RESID_LN) {
                epilogue_resid_ln<CG>(p, tm_out0, tm_out1, tm_out2, box, ln_xch, t_base, row0, q, sub, lane, BN, oscale,
                                      &tmem_empty_bar[acc], xr);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
                continue;
            }
            if (sub >= nck) release(&tmem_empty_bar[acc]);            // a narrow tile leaves this warp without a chunk
            for (int c = sub; c < nck; c += EPI_SUBS) {
                const int c0 = c * CW;
                uint32_t r[16];
                tmem_ld_32x16(t_base + (uint32_t)c0, r);
                const int col = n_tile * BN + c0;
                float4 bv[4];
                load_vec16(bv, p.bias + col);
                if (lane == 0) GEMM_BULK_WAIT_READ();             // the store that last used the box has read it
                tmem_ld_wait();
                if (c + EPI_SUBS >= nck) release(&tmem_empty_bar[acc]);
                __syncwarp();
                float2 v[8];                                          // packed pairs: FFMA2 (common.cuh)
                const float2 osc2 = splat2(oscale);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[2 * i] = ffma2(pair_of(r, 2 * i), osc2, make_float2(bv[i].x, bv[i].y));
                    v[2 * i + 1] = ffma2(pair_of(r, 2 * i + 1), osc2, make_float2(bv[i].z, bv[i].w));
                }
                if (EPI == EPI_F32 || EPI == EPI_RESID) {
                    // one box of 32 rows x 64 B
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        sts128(box_s + box_off_f32(lane, i), make_float4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y));
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (EPI == EPI_RESID) tma_reduce_add_2d(&tm_out0, box, col, row0);
                        else tma_store_2d(&tm_out0, box, col, row0);
                        bulk_commit_group();
                    }
                } else {
                    // fp16 hi / lo outputs: two boxes of 32 rows x 32 B
                    if (EPI == EPI_GELU_SPLIT) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = gelu_erf_fast2(v[e]);
                    }
                    stage_split16(box_s, lane, v);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (EPI == EPI_GELU_SPLIT) {
                            tma_store_2d(&tm_out0, box, col, row0);
                            tma_store_2d(&tm_out1, box + 1024, col, row0);
                        } else {                                      // head planes: 16 columns never straddle a plane (hds % 16 == 0)
                            tma_store_3d(&tm_out0, box, col % p.hds, row0, col / p.hds);
                            tma_store_3d(&tm_out1, box + 1024, col % p.hds, row0, col / p.hds);
                        }
                        bulk_commit_group();
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        GEMM_TR(tr_work)
#ifdef PAFUSE_GEMM_TRACE
        if (blockIdx.x == 0 && lane == 0 && q == 0 && sub == 0 && walk.count > 0)
            printf("gemm_trace EPI=%d WRES=%d N=%d K=%d tiles=%d cycles/tile: wait_tmem_full %lld | epilogue_work %lld | of_which_staging_box_wait %lld | LN pass1 %lld merge1 %lld pass2 %lld merge2 %lld pass3 %lld\n", EPI,
                   (int)WRES, p.N, p.K, walk.count, tr_wait / walk.count, tr_work / walk.count, g_gemm_tr_bulk / walk.count,
                   g_gemm_tr_ln[0] / walk.count, g_gemm_tr_ln[1] / walk.count, g_gemm_tr_ln[2] / walk.count, g_gemm_tr_ln[3] / walk.count,
                   g_gemm_tr_ln[4] / walk.count);
#endif
        if (lane == 0) bulk_wait_group<0>();                          // all output boxes are written before the CTA retires
    }

    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();            // nobody leaves while the peer may still signal it
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<CG>(tmem_base, TMEM_COLS);
    }
}


// ====================================================================================================
// Fused MLP (mixste.py:37-43 + the residual add and LayerNorms of EPI_RESID_LN):
//     x <- x + fc2(GELU(fc1(a)))  [+ the norms that follow],  a = fp16 hi/lo LayerNorm output, C <= 256.
// The hidden activations [M, 2C] never leave the SM: per 128-column hidden chunk j
//     acc1 = a W1[j]^T            (G1: A from shared memory, resident for the whole 256-row tile)
//     h    = GELU(acc1 + b1)      (epilogue warps: TMEM -> registers -> fp16 hi/lo -> TMEM, in place over acc1)
//     acc2 += h W2[:, j]^T        (G2: A operand read from tensor memory, like P in the attention kernel)
// and the EPI_RESID_LN epilogue runs on acc2.  Against the fc1 + fc2 launches this removes the write and the
// read of h (16 of the 32 bytes per element-row the MLP moved through DRAM) and one launch.
// Tensor memory (512 columns): acc2 [0, C); two chunk buffers of 128 columns at 256 / 384, each holding acc1 (fp32)
// and then h (hi in the first hc/2 columns, lo in the next hc/2, packed fp16 pairs).  Shared memory: the A tile
// (C/64 boxes x hi/lo x 16 KB), a ring of 16 KB weight stages (W1: one 64-wide K box of this CTA's 64 hidden rows,
// hi + lo; W2: 32 K columns of this CTA's C/2 output rows, hi + lo), the epilogue staging boxes.
// Issue order per tile (one thread of the even CTA): G1(0) G1(1) | G2(j) G1(j+2) ... | G2(n-2) G2(n-1): the tensor
// pipe executes in issue order, so G1(j+2) overwrites a chunk buffer only after G2(j) has read h from it, and it
// works on G1(j+1) while the epilogue warps turn chunk j into h.
// The first version used 64-column chunks: its G1 MMAs (N = 64, 32 cycles of math) form a chain of dependent
// accumulations into one tile and ran at the pipe's latency (~160 cycles per MMA measured, profiles/r1n_*): the
// kernel was 20-28 % SLOWER than the two GEMM launches.
constexpr int MLP_HC = 128;                      // hidden columns per chunk (the last chunk may be 64)
constexpr int MLP_STAGE_BYTES = 16384;
constexpr int MLP_A_BOX_BYTES = BM * BKW * 2;    // 16 KB: 128 rows x 64 fp16
constexpr int MLP_TM_BUF = 256;                  // chunk buffer b at MLP_TM_BUF + 128 b

struct MlpParams {
    long long M;
    int C, hidden, nch;              // channels, hidden width (2C), hidden chunks
    int m_tiles;                     // tiles of 256 rows
    int stages;                      // weight ring depth
    int kb;                          // 64-wide K boxes of the A tile = ceil(C / 64)
    float out_scale;
    const float* bias1;              // [2C]
    KernelParams ep;                 // the EPI_RESID_LN epilogue's view: N = C, bias = b2, ln
};

// D[tmem] (+)= A[tmem] * B[smem desc], CTA pair
__device__ __forceinline__ void umma_f16_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 accumulator columns -> bias + GELU -> 8 packed fp16 hi pairs and 8 lo pairs
__device__ __forceinline__ void mlp_gelu_split16(const uint32_t (&r)[16], const float* bias, float oscale, uint32_t (&hi)[8],
                                                 uint32_t (&lo)[8]) {
    float4 bv[4];
    load_vec16(bv, bias);
    const float2 osc2 = splat2(oscale);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 va = gelu_erf_fast2(ffma2(pair_of(r, 2 * e), osc2, make_float2(bv[e].x, bv[e].y)));
        const float2 vb = gelu_erf_fast2(ffma2(pair_of(r, 2 * e + 1), osc2, make_float2(bv[e].z, bv[e].w)));
        split_pair_sat2(va, hi[2 * e], lo[2 * e]);
        split_pair_sat2(vb, hi[2 * e + 1], lo[2 * e + 1]);
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                 const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                 const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_o_hi,
                 const __grid_constant__ CUtensorMap tm_o_lo, const MlpParams p) {
    constexpr int CG = 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ float ln_xch[128 * EPI_SUBS];
    __shared__ __align__(8) uint64_t w_full[MAX_STAGES], w_empty[MAX_STAGES];
    __shared__ __align__(8) uint64_t a_full, a_empty, acc2_full, acc2_empty;
    __shared__ __align__(8) uint64_t acc1_full[2], h_full[2];
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = smem_raw;
    uint8_t* smem_a = smem;                                           // [kb][hi, lo][128 rows x 128 B]
    uint8_t* smem_w = smem_a + (size_t)p.kb * 2 * MLP_A_BOX_BYTES;    // weight stage ring
    uint8_t* smem_box = smem_w + (size_t)p.stages * MLP_STAGE_BYTES;  // epilogue staging, 4 KB per warp

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int C = p.C, nch = p.nch;
    const int pair = (int)blockIdx.x / CG, num_pairs = (int)gridDim.x / CG;
    const int n_local = pair < p.m_tiles ? (p.m_tiles - pair + num_pairs - 1) / num_pairs : 0;
    const int w2_rows = C / CG;                                       // output rows of W2 this CTA loads
    auto chunk_width = [&](int j) { return p.hidden - j * MLP_HC >= MLP_HC ? MLP_HC : p.hidden - j * MLP_HC; };

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        if (smem_u32(smem_raw) & 1023u) {
            printf("pafuse: dynamic shared memory is not 1024-byte aligned\n");
            asm volatile("trap;");
        }
        prefetch_tensormap(&tm_a_hi); prefetch_tensormap(&tm_a_lo);
        prefetch_tensormap(&tm_w1_hi); prefetch_tensormap(&tm_w1_lo);
        prefetch_tensormap(&tm_w2_hi); prefetch_tensormap(&tm_w2_lo);
        prefetch_tensormap(&tm_x); prefetch_tensormap(&tm_o_hi); prefetch_tensormap(&tm_o_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&w_full[s], 1);                                 // the leader's arrive.expect_tx
            mbar_init(&w_empty[s], 1);                                // one tcgen05.commit
        }
        mbar_init(&a_full, 1);
        mbar_init(&a_empty, 1);
        mbar_init(&acc2_full, 1);
        mbar_init(&acc2_empty, EPI_WARPS * CG);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc1_full[b], 1);
            mbar_init(&h_full[b], EPI_WARPS * CG);                    // every epilogue warp of the pair has written its part of h
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<CG>(&tmem_base_slot, TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    pdl_wait();

    if (warp < EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " PAFUSE_REGS_CTRL ";");
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0 && n_local > 0) {
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&]() {
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            const uint32_t w2_bytes = (uint32_t)(2 * w2_rows * 64);
            // (Start offsets per CTA pair, to keep the pairs from running their DRAM-heavy epilogues in lock-step, and
            // weight rings of 2 / 3 / 4 stages all measured the same: profiles/r1n_mlp_fused.txt.)
            for (int i = 0; i < n_local; ++i) {
                const int tile = pair + i * num_pairs;
                const int m0 = (tile * CG + (int)cta_rank) * BM;
                mbar_wait(&a_empty, (uint32_t)(i & 1) ^ 1);           // every G1 of the previous tile has retired
                if (leader) mbar_arrive_expect_tx(&a_full, (uint32_t)(2 * p.kb * 2 * MLP_A_BOX_BYTES));
                for (int kb = 0; kb < p.kb; ++kb) {
                    tma_load_2d_pair(smem_a + (size_t)(2 * kb) * MLP_A_BOX_BYTES, &tm_a_hi, &a_full, kb * BKW, m0);
                    tma_load_2d_pair(smem_a + (size_t)(2 * kb + 1) * MLP_A_BOX_BYTES, &tm_a_lo, &a_full, kb * BKW, m0);
                }
                for (int s = 0; s < nch + 2; ++s) {
                    if (s >= 2) {                                     // W2 of chunk s - 2: 32 K columns per stage
                        const int j = s - 2;
                        const int nst = chunk_width(j) / 32;
                        for (int h2 = 0; h2 < nst; ++h2) {
                            mbar_wait(&w_empty[stage], phase ^ 1);
                            uint8_t* st = smem_w + (size_t)stage * MLP_STAGE_BYTES;
                            if (leader) mbar_arrive_expect_tx(&w_full[stage], 2 * w2_bytes);
                            tma_load_2d_pair(st, &tm_w2_hi, &w_full[stage], j * MLP_HC + 32 * h2, (int)cta_rank * w2_rows);
                            tma_load_2d_pair(st + 8192, &tm_w2_lo, &w_full[stage], j * MLP_HC + 32 * h2, (int)cta_rank * w2_rows);
                            advance();
                        }
                    }
                    if (s < nch) {
                        // W1 of chunk s: this CTA's half of the chunk's rows, one 64-wide K box per stage.  The box is
                        // always 64 rows; a 64-column last chunk uses its first 32 (the rest is the peer's half, or
                        // zero-fill past the tensor).
                        const int n0 = s * MLP_HC + (int)cta_rank * (chunk_width(s) / CG);
                        for (int kb = 0; kb < p.kb; ++kb) {
                            mbar_wait(&w_empty[stage], phase ^ 1);
                            uint8_t* st = smem_w + (size_t)stage * MLP_STAGE_BYTES;
                            if (leader) mbar_arrive_expect_tx(&w_full[stage], (uint32_t)(2 * 2 * 8192));
                            tma_load_2d_pair(st, &tm_w1_hi, &w_full[stage], kb * BKW, n0);
                            tma_load_2d_pair(st + 8192, &tm_w1_lo, &w_full[stage], kb * BKW, n0);
                            advance();
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the even CTA) =====================
        if (lane == 0 && leader && n_local > 0) {
            const uint32_t idesc2 = make_idesc_f16((uint32_t)(BM * CG), (uint32_t)C);
            const uint32_t sa = smem_u32(smem_a), sw = smem_u32(smem_w);
            const uint32_t d_acc2 = tmem_base;
            const uint64_t a_desc0 = make_smem_desc_sw128(sa);        // + byte offset / 16 selects box, half and K slice
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&]() {
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            uint32_t g1 = 0, g2 = 0;                                  // chunks issued so far (buffer = count & 1)
            for (int i = 0; i < n_local; ++i) {
                mbar_wait(&a_full, (uint32_t)(i & 1));                // the A tile landed (in both CTAs)
                tcgen05_fence_after();
                for (int s = 0; s < nch + 2; ++s) {
                    if (s >= 2) {
                        // ---- G2(j): acc2 += h W2[:, j]^T, h read from the chunk buffer in tensor memory
                        const int j = s - 2;
                        const int hc = chunk_width(j);
                        const uint32_t b = g2 & 1u, use = g2 >> 1;
                        mbar_wait(&h_full[b], use & 1u);              // h of this chunk is in tensor memory (both CTAs)
                        if (j == 0) mbar_wait(&acc2_empty, (uint32_t)(i & 1) ^ 1u);   // the previous tile's epilogue is done with acc2
                        tcgen05_fence_after();
                        const uint32_t h_hi = tmem_base + (uint32_t)MLP_TM_BUF + b * 128u, h_lo = h_hi + (uint32_t)(hc / 2);
                        for (int h2 = 0; h2 < hc / 32; ++h2) {
                            mbar_wait(&w_full[stage], phase);
                            tcgen05_fence_after();
                            const uint32_t st = sw + (uint32_t)stage * MLP_STAGE_BYTES;
                            const uint64_t dwh0 = make_smem_desc(st, 512u, 4u), dwl0 = make_smem_desc(st + 8192u, 512u, 4u);   // 64-byte swizzle
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint32_t kk = (uint32_t)(2 * h2 + k);                  // 16-wide K step inside the chunk
                                const uint64_t ko = (uint64_t)(k * 2);                       // 32 bytes, in 16-byte units
                                umma_f16_ts2(d_acc2, h_lo + 8u * kk, dwh0 + ko, idesc2, (j | (int)kk) != 0 ? 1u : 0u);
                                umma_f16_ts2(d_acc2, h_hi + 8u * kk, dwl0 + ko, idesc2, 1u);
                                umma_f16_ts2(d_acc2, h_hi + 8u * kk, dwh0 + ko, idesc2, 1u);
                            }
                            umma_commit<CG>(&w_empty[stage]);
                            advance();
                        }
                        if (j == nch - 1) umma_commit<CG>(&acc2_full);
                        ++g2;
                    }
                    if (s < nch) {
                        // ---- G1(s): chunk buffer <- A W1[s]^T.  The buffer held h of chunk s - 2, whose G2 was issued
                        // above / earlier by this thread: the tensor pipe executes in issue order.
                        const int hc = chunk_width(s);
                        const uint32_t idesc1 = make_idesc_f16((uint32_t)(BM * CG), (uint32_t)hc);
                        const uint32_t b = g1 & 1u;
                        const uint32_t d = tmem_base + (uint32_t)MLP_TM_BUF + b * 128u;
                        for (int kb = 0; kb < p.kb; ++kb) {
                            mbar_wait(&w_full[stage], phase);
                            tcgen05_fence_after();
                            const uint32_t st = sw + (uint32_t)stage * MLP_STAGE_BYTES;
                            const int k_left = C - kb * BKW;
                            const int nk = k_left >= BKW ? BKW / UK : (k_left + UK - 1) / UK;
                            const uint64_t dah0 = a_desc0 + (uint64_t)((2 * kb) * (MLP_A_BOX_BYTES >> 4));
                            const uint64_t dal0 = dah0 + (uint64_t)(MLP_A_BOX_BYTES >> 4);
                            const uint64_t dwh0 = make_smem_desc_sw128(st), dwl0 = make_smem_desc_sw128(st + 8192u);
#pragma unroll
                            for (int k = 0; k < BKW / UK; ++k) {
                                if (k < nk) {
                                    const uint64_t ko = (uint64_t)(k * 2);                   // 32 bytes along K, in 16-byte units
                                    umma_f16_ss<CG>(d, dal0 + ko, dwh0 + ko, idesc1, (kb | k) != 0 ? 1u : 0u);
                                    umma_f16_ss<CG>(d, dah0 + ko, dwl0 + ko, idesc1, 1u);
                                    umma_f16_ss<CG>(d, dah0 + ko, dwh0 + ko, idesc1, 1u);
                                }
                            }
                            umma_commit<CG>(&w_empty[stage]);
                            advance();
                        }
                        umma_commit<CG>(&acc1_full[b]);
                        if (s == nch - 1) umma_commit<CG>(&a_empty);  // the A tile may be overwritten
                        ++g1;
                    }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " PAFUSE_REGS_EPI ";");
        // ===================== epilogue warps: GELU per chunk, then residual + LayerNorms per tile =====================
        const int q = warp & 3;                                       // TMEM lane quarter == warp % 4
        const int sub = (warp - EPI_WARP0) >> 2;                      // which quarter of the chunk's columns
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const float oscale = p.out_scale;
        uint8_t* box = smem_box + (warp - EPI_WARP0) * STG_WARP_BYTES;
        uint32_t g = 0;
        for (int i = 0; i < n_local; ++i) {
            const int tile = pair + i * num_pairs;
            const int row0 = (tile * CG + (int)cta_rank) * BM + q * 32;
            for (int j = 0; j < nch; ++j, ++g) {
                const int hc = chunk_width(j);
                const uint32_t b = g & 1u, use = g >> 1;
                const uint32_t buf = tmem_base + lane_sel + (uint32_t)MLP_TM_BUF + b * 128u;
                const int wcols = hc / EPI_SUBS;                      // this warp's columns of the chunk: 32, or 16 for a 64-wide chunk
                const int c_first = sub * wcols;
                const float* bias = p.bias1 + j * MLP_HC + c_first;
                mbar_wait(&acc1_full[b], use & 1u);
                tcgen05_fence_after();
                uint32_t r0[16], r1[16];
                tmem_ld_32x16(buf + (uint32_t)c_first, r0);
                if (hc == MLP_HC) tmem_ld_32x16(buf + (uint32_t)c_first + 16u, r1);
                tmem_ld_wait();
                // h goes over acc1 in place, and this warp's hi / lo columns are other warps' accumulator columns:
                // all warps of the lane quarter must have read before any of them writes
                quarter_bar_sync(1 + q);
                uint32_t hi[8], lo[8];
                const uint32_t hi_addr = buf + (uint32_t)(c_first / 2), lo_addr = hi_addr + (uint32_t)(hc / 2);
                mlp_gelu_split16(r0, bias, oscale, hi, lo);
                tmem_st_32x8(hi_addr, hi);
                tmem_st_32x8(lo_addr, lo);
                if (hc == MLP_HC) {
                    mlp_gelu_split16(r1, bias + 16, oscale, hi, lo);
                    tmem_st_32x8(hi_addr + 8u, hi);
                    tmem_st_32x8(lo_addr + 8u, lo);
                }
                tmem_st_wait_all();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&h_full[b], 0);
            }
            float4 xr[4];
            if (sub * CW < C) ln_fetch_x(p.ep, xr, row0, sub * CW, lane);   // in flight while the last G2 finishes
            mbar_wait(&acc2_full, (uint32_t)(i & 1));
            tcgen05_fence_after();
            epilogue_resid_ln<CG>(p.ep, tm_x, tm_o_hi, tm_o_lo, box, ln_xch, tmem_base + lane_sel, row0, q, sub, lane, C,
                                  oscale, &acc2_empty, xr);
        }
        if (lane == 0) bulk_wait_group<0>();
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<CG>(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

// operand boxes: box_k (64 / 32) columns x box_rows rows, swizzle span = the bytes of one box row
int make_map_f16(CUtensorMap* map, const void* ptr, long long rows, int K, int box_rows, int box_k = BKW) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%d box_rows=%d ptr=%p", (int)r, rows, K, box_rows,
                       ptr);
        return -2;
    }
    return 0;
}

// output boxes of the epilogue: CW = 16 columns x 32 rows, swizzle = the 64 (fp32) / 32 (fp16) bytes of one box row
int make_map_out(CUtensorMap* map, const void* ptr, long long rows, int N, bool f32) {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)N * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)CW, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                          const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          f32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(out) failed (%d) rows=%lld N=%d ptr=%p", (int)r, rows, N, ptr);
        return -2;
    }
    return 0;
}

// head planes [24][rows_cap][hds] fp16: box = CW columns x 32 rows of one plane
int make_map_planes(CUtensorMap* map, const void* ptr, long long rows_cap, int hds) {
    cuuint64_t dims[3] = {(cuuint64_t)hds, (cuuint64_t)rows_cap, 24};
    cuuint64_t strides[2] = {(cuuint64_t)hds * 2, (cuuint64_t)rows_cap * hds * 2};
    cuuint32_t box[3] = {(cuuint32_t)CW, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(planes) failed (%d) rows_cap=%lld hds=%d ptr=%p", (int)r, rows_cap, hds, ptr);
        return -2;
    }
    return 0;
}

int g_force_cg = 0;      // PAFUSE_GEMM_CTA_GROUP=1|2 overrides the default (2)
int g_wres_enabled = 1;  // PAFUSE_GEMM_WRES=0 disables the weight-stationary mode
int g_wres_min_stages = 4;  // with 3 stages (C = 384) the resident mode was slower than streaming (round-1 session measurement, 2-3 % on the body GEMMs; no capture kept)

template <int EPI, int CG, bool WRES>
int launch_epi(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
               const CUtensorMap& o0, const CUtensorMap& o1, const CUtensorMap& o2, const KernelParams& kp, int grid,
               int smem, cudaStream_t st) {
    auto kern = gemm_f16x3_kernel<EPI, CG, WRES>;
    static bool configured[MAX_DEVICES] = {false};                    // per template instance and per device
    const int dev = current_device_slot();
    if (!configured[dev]) {
        cudaFuncAttributes fa;
        PAFUSE_CUDA_OK(cudaFuncGetAttributes(&fa, kern));              // static shared memory counts against the 227 KiB
        const int static_window = ((int)fa.sharedSizeBytes + 1023) / 1024 * 1024;
        if (static_window > (EPI == EPI_RESID_LN ? SMEM_SLACK_LN : SMEM_SLACK)) {
            set_last_error("gemm: kernel has %d bytes of static shared memory, the ring was sized for less", (int)fa.sharedSizeBytes);
            return -1;
        }
        PAFUSE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - static_window));
        configured[dev] = true;
    }
    PAFUSE_CUDA_OK(launch_chain(kern, dim3((unsigned)grid), dim3(NUM_THREADS), (size_t)smem, st, CG, ah, al, wh, wl, o0, o1,
                                o2, kp));
    PAFUSE_LAUNCH_OK();
    return 0;
}

template <int EPI, int CG>
int launch_mode(bool wres, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
                const CUtensorMap& o0, const CUtensorMap& o1, const CUtensorMap& o2, const KernelParams& kp, int grid,
                int smem, cudaStream_t st) {
    return wres ? launch_epi<EPI, CG, true>(ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st)
                : launch_epi<EPI, CG, false>(ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st);
}

template <int CG>
int launch_cg(const GemmArgs& g, cudaStream_t st) {
    const int BN = gemm_pick_block_n(g.N);
    if (BN < 32 || BN % 32 != 0 || g.K % 8 != 0 || g.N % CW != 0) {
        set_last_error("gemm: unsupported shape N=%d K=%d (block_n=%d)", g.N, g.K, BN);
        return -1;
    }
    KernelParams kp;
    kp.M = g.M;
    kp.N = g.N;
    kp.K = g.K;
    kp.block_n = BN;
    kp.m_tiles = (int)((g.M + BM * CG - 1) / (BM * CG));
    kp.n_tiles = g.N / BN;
    kp.out_scale = g.out_scale;
    kp.bias = g.bias;
    kp.hds = g.planes.hds;
    kp.ln = g.ln;
#ifdef PAFUSE_ABLATE
    kp.ablate = getenv("PAFUSE_ABLATE") ? atoi(getenv("PAFUSE_ABLATE")) : 0;
#endif
    if (g.epilogue == EPI_RESID_LN) {
        if (!gemm_can_fuse_ln(g.N) || kp.n_tiles != 1 || !g.ln.x || g.ln.x != g.out_f32 || !g.ln.g1 || !g.ln.b1 ||
            !g.out_hi || !g.out_lo || (g.ln.g0 && !g.ln.b0)) {
            set_last_error("gemm: EPI_RESID_LN needs N <= 256 in one tile, x == out_f32 and the norm parameters (N=%d)", g.N);
            return -1;
        }
    }
    const int num_sms = device_sm_count();
    const int sms = g.sm_limit > 0 && g.sm_limit < num_sms ? g.sm_limit : num_sms;
    const int max_groups = sms / CG > 0 ? sms / CG : 1;
    const int WN = BN / CG;
    const int slack = g.epilogue == EPI_RESID_LN ? SMEM_SLACK_LN : SMEM_SLACK;

    // Weight-stationary mode: the W slice of one n tile (all of K, hi and lo) stays in shared memory and only A
    // is streamed, in 32-wide K stages.  Chosen when the slice leaves room for >= 3 stages; the K = 2C layers
    // (fc2) do not fit and stream both operands.
    bool wres = false;
    kp.a_bk = BKW;
    kp.w_res_bytes = 0;
    kp.slots = 0;
    if (g_wres_enabled && g.K % 32 == 0 && kp.n_tiles <= max_groups) {
        const int w_res = ((g.K + BKW - 1) / BKW) * 2 * WN * BKW * 2;
        const int a_stage = 2 * BM * 32 * 2;
        const int stages = (SMEM_LIMIT - slack - STG_BYTES - w_res) / a_stage;
        if (stages >= g_wres_min_stages) {
            wres = true;
            kp.a_bk = 32;
            kp.w_res_bytes = w_res;
            kp.stage_bytes = a_stage;
            kp.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
            kp.slots = max_groups / kp.n_tiles;
            if (kp.slots > kp.m_tiles) kp.slots = kp.m_tiles;
        }
    }
    if (!wres) {
        kp.stage_bytes = 2 * BM * BKW * 2 + 2 * WN * BKW * 2;
        int stages = (SMEM_LIMIT - slack - STG_BYTES) / kp.stage_bytes;
        kp.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    }

    CUtensorMap ah, al, wh, wl;
    if (int rc = make_map_f16(&ah, g.a_hi, g.M, g.K, BM, kp.a_bk)) return rc;
    if (int rc = make_map_f16(&al, g.a_lo, g.M, g.K, BM, kp.a_bk)) return rc;
    if (int rc = make_map_f16(&wh, g.w_hi, g.N, g.K, WN)) return rc;
    if (int rc = make_map_f16(&wl, g.w_lo, g.N, g.K, WN)) return rc;
    CUtensorMap o0, o1, o2;
    if (g.epilogue == EPI_RESID_LN) {
        if (int rc = make_map_out(&o0, g.out_f32, g.M, g.N, true)) return rc;
        if (int rc = make_map_out(&o1, g.out_hi, g.M, g.N, false)) return rc;
        if (int rc = make_map_out(&o2, g.out_lo, g.M, g.N, false)) return rc;
    } else if (g.epilogue == EPI_PLANES) {
        if (g.planes.hds < 16 || g.planes.hds % 16 != 0 || g.N != 24 * g.planes.hds || g.planes.rows_cap < g.M) {
            set_last_error("gemm: bad head planes (hds=%d N=%d rows_cap=%lld M=%lld)", g.planes.hds, g.N, g.planes.rows_cap, g.M);
            return -1;
        }
        if (int rc = make_map_planes(&o0, g.planes.hi, g.planes.rows_cap, g.planes.hds)) return rc;
        if (int rc = make_map_planes(&o1, g.planes.lo, g.planes.rows_cap, g.planes.hds)) return rc;
    } else if (g.epilogue == EPI_GELU_SPLIT) {
        if (int rc = make_map_out(&o0, g.out_hi, g.M, g.N, false)) return rc;
        if (int rc = make_map_out(&o1, g.out_lo, g.M, g.N, false)) return rc;
    } else {
        if (int rc = make_map_out(&o0, g.out_f32, g.M, g.N, true)) return rc;
        o1 = o0;
    }
    if (g.epilogue != EPI_RESID_LN) o2 = o0;
    const int smem = kp.w_res_bytes + kp.stages * kp.stage_bytes + STG_BYTES;
    int grid;
    if (wres) {
        grid = kp.slots * kp.n_tiles * CG;
    } else {
        const long long tiles = (long long)kp.m_tiles * kp.n_tiles;
        grid = (int)(tiles < max_groups ? tiles : max_groups) * CG;
    }
    switch (g.epilogue) {
        case EPI_F32: return launch_mode<EPI_F32, CG>(wres, ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st);
        case EPI_GELU_SPLIT: return launch_mode<EPI_GELU_SPLIT, CG>(wres, ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st);
        case EPI_RESID: return launch_mode<EPI_RESID, CG>(wres, ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st);
        case EPI_PLANES: return launch_mode<EPI_PLANES, CG>(wres, ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st);
        case EPI_RESID_LN: return launch_mode<EPI_RESID_LN, CG>(wres, ah, al, wh, wl, o0, o1, o2, kp, grid, smem, st);
    }
    set_last_error("gemm: bad epilogue %d", g.epilogue);
    return -1;
}


int launch_mlp(const MlpArgs& g, cudaStream_t st) {
    const int C = g.C, hidden = 2 * C;
    if (!mlp_can_fuse(C) || !g.x || !g.out_hi || !g.out_lo || !g.ln.g1 || !g.ln.b1 || (g.ln.g0 && !g.ln.b0) || g.ln.x != g.x) {
        set_last_error("mlp_fused: needs C %% 32 == 0, 64 <= C <= 256, the residual stream and the norm parameters (C=%d)", C);
        return -1;
    }
    MlpParams mp;
    mp.M = g.M;
    mp.C = C;
    mp.hidden = hidden;
    mp.nch = (hidden + MLP_HC - 1) / MLP_HC;
    mp.m_tiles = (int)((g.M + 2 * BM - 1) / (2 * BM));
    mp.kb = (C + BKW - 1) / BKW;
    mp.out_scale = g.out_scale;
    mp.bias1 = g.b1;
    mp.ep = KernelParams();
    mp.ep.M = g.M;
    mp.ep.N = C;
    mp.ep.K = hidden;
    mp.ep.block_n = C;
    mp.ep.out_scale = g.out_scale;
    mp.ep.bias = g.b2;
    mp.ep.ln = g.ln;

    auto kern = mlp_fused_kernel;
    static int max_dyn_dev[MAX_DEVICES] = {0};                         // per device (function attributes are per device)
    const int dev = current_device_slot();
    if (!max_dyn_dev[dev]) {
        cudaFuncAttributes fa;
        PAFUSE_CUDA_OK(cudaFuncGetAttributes(&fa, kern));
        const int static_window = ((int)fa.sharedSizeBytes + 1023) / 1024 * 1024;
        PAFUSE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - static_window));
        max_dyn_dev[dev] = SMEM_LIMIT - static_window;
    }
    const int max_dyn = max_dyn_dev[dev];
    const int fixed = mp.kb * 2 * MLP_A_BOX_BYTES + STG_BYTES;
    int stages = (max_dyn - fixed) / MLP_STAGE_BYTES;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) {
        set_last_error("mlp_fused: C=%d leaves room for %d weight stages only", C, stages);
        return -1;
    }
    mp.stages = stages;
    const int smem = fixed + stages * MLP_STAGE_BYTES;

    CUtensorMap ah, al, w1h, w1l, w2h, w2l, ox, oh, ol;
    if (int rc = make_map_f16(&ah, g.a_hi, g.M, C, BM)) return rc;
    if (int rc = make_map_f16(&al, g.a_lo, g.M, C, BM)) return rc;
    if (int rc = make_map_f16(&w1h, g.w1_hi, hidden, C, MLP_HC / 2)) return rc;     // 64-row boxes (this CTA's half of a chunk)
    if (int rc = make_map_f16(&w1l, g.w1_lo, hidden, C, MLP_HC / 2)) return rc;
    if (int rc = make_map_f16(&w2h, g.w2_hi, C, hidden, C / 2, 32)) return rc;
    if (int rc = make_map_f16(&w2l, g.w2_lo, C, hidden, C / 2, 32)) return rc;
    if (int rc = make_map_out(&ox, g.x, g.M, C, true)) return rc;
    if (int rc = make_map_out(&oh, g.out_hi, g.M, C, false)) return rc;
    if (int rc = make_map_out(&ol, g.out_lo, g.M, C, false)) return rc;
    const int num_sms = device_sm_count();
    const int sms = g.sm_limit > 0 && g.sm_limit < num_sms ? g.sm_limit : num_sms;
    const int max_pairs = sms / 2 > 0 ? sms / 2 : 1;
    const int grid = 2 * (mp.m_tiles < max_pairs ? mp.m_tiles : max_pairs);
    PAFUSE_CUDA_OK(launch_chain(kern, dim3((unsigned)grid), dim3(NUM_THREADS), (size_t)smem, st, 2, ah, al, w1h, w1l, w2h, w2l,
                                ox, oh, ol, mp));
    PAFUSE_LAUNCH_OK();
    return 0;
}

}  // namespace

bool mlp_can_fuse(int C) { return C % 32 == 0 && C >= 64 && C <= 256; }

int launch_mlp_fused(const MlpArgs& g, cudaStream_t st) {
    if (g.M == 0) return 0;
    if (int rc = gemm_init()) return rc;
    return launch_mlp(g, st);
}

namespace {

}  // namespace

int gemm_pick_block_n(int N) {
    // largest UMMA-legal N (multiple of 32 here, <= 256) that divides the layer width; the
    // path's widths are 224/256/384/448/512/672/768/1152 -> 224/256/192/224/256/224/256/192
    for (int bn = 256; bn >= 32; bn -= 32)
        if (N % bn == 0) return bn;
    return 0;
}

bool gemm_can_fuse_ln(int N) { return N % 32 == 0 && N <= 256 && gemm_pick_block_n(N) == N; }

int gemm_init() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PAFUSE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        set_last_error("cuTensorMapEncodeTiled not available from the driver");
        return -2;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const char* e = getenv("PAFUSE_GEMM_CTA_GROUP");
    g_force_cg = e ? atoi(e) : 0;
    if (const char* w = getenv("PAFUSE_GEMM_WRES")) g_wres_enabled = atoi(w);
    if (const char* w = getenv("PAFUSE_GEMM_WRES_MIN_STAGES")) g_wres_min_stages = atoi(w);
    return 0;
}

void gemm_set_cta_group(int cg) { g_force_cg = cg; }
void gemm_set_weight_stationary(int on) { g_wres_enabled = on; }

int launch_gemm_tcgen05(const GemmArgs& g, cudaStream_t st) {
    if (g.M == 0) return 0;
    if (int rc = gemm_init()) return rc;
    return g_force_cg == 1 ? launch_cg<1>(g, st) : launch_cg<2>(g, st);
}

// ------------------------------------------------------------------ debug reference on CUDA cores
// Same contract, evaluated as (A_hi+A_lo)*(W_hi+W_lo) in fp32 FMAs: used by the unit
// tests to check the tensor-core kernel on device and selectable for bisecting.
// Not a production path.
__global__ void gemm_simt_kernel(GemmArgs g) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.M * g.N) return;
    long long m = idx / g.N;
    int n = (int)(idx % g.N);
    const op_t* ah = g.a_hi + (size_t)m * g.K;
    const op_t* al = g.a_lo + (size_t)m * g.K;
    const op_t* wh = g.w_hi + (size_t)n * g.K;
    const op_t* wl = g.w_lo + (size_t)n * g.K;
    float acc = 0.f;
    for (int k = 0; k < g.K; ++k) acc = fmaf(join_op(ah[k], al[k]), join_op(wh[k], wl[k]), acc);
    float v = fmaf(acc, g.out_scale, g.bias[n]);
    size_t o = (size_t)m * g.N + n;
    if (g.epilogue == EPI_PLANES) {
        op_t h, l;
        split_op(v, h, l);
        size_t po = ((size_t)(n / g.planes.hds) * g.planes.rows_cap + m) * g.planes.hds + n % g.planes.hds;
        g.planes.hi[po] = h;
        g.planes.lo[po] = l;
    } else if (g.epilogue == EPI_F32) {
        g.out_f32[o] = v;
    } else if (g.epilogue == EPI_RESID) {
        g.out_f32[o] += v;
    } else {
        op_t h, l;
        split_op(gelu_erf(v), h, l);
        g.out_hi[o] = h;
        g.out_lo[o] = l;
    }
}

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
    long long total = g.M * g.N;
    if (total == 0) return 0;
    gemm_simt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g);
    PAFUSE_LAUNCH_OK();
    return 0;
}

}  // namespace pafuse
