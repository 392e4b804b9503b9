// Short-sequence attention of the STE / TTE blocks (mixste.py:63-82, comb=False) on the
// sm_100a tensor cores.
//
// The sequences are tiny (24 / 68 / 42 joints, 27 frames) while tcgen05.mma wants M = 128, so a
// tile packs G whole sequences ("groups"), each padded to Lp = L rounded up to 32 rows:
// body 4 x (24 -> 32), face 1 x (68 -> 96), hands 2 x (42 -> 64) joints of consecutive (clip, hypothesis,
// frame) sequences; 4 joints x (27 -> 32) frames of one clip-hypothesis for the temporal blocks.  The
// padding costs nothing in HBM: the TMA box is Lp rows long in a tensor dimension of extent L, the
// rows past L are zero-filled by the TMA unit.  With 32-row alignment every softmax warp (32 tile rows
// = 32 TMEM lanes) lies inside ONE group, so the score columns it needs are whole 32-column chunks
// with compile-time masks (the dense packing of the previous version made every row read and
// exponentiate up to 96 columns for 24 live keys, with run-time masks: 2-3 K instructions per unit).
//
// Per unit = (tile, head): S = Q K^T (128 x 128, three f16x3 passes into TMEM); softmax over the
// row's own group; the un-normalised probabilities go back to tensor memory as fp16 hi/lo OVER the
// scores (zeros outside the group; the A operand of the second MMA is read from TMEM); O = P V with V
// as an MN-major shared-memory operand.  The off-diagonal blocks are wasted tensor math, which is cheap.
//
// Inputs are the per-head planes the qkv GEMM epilogue writes (EPI_PLANES): fp16 hi/lo arrays
// [which(q,k,v) * 8 + head][token][hds], hds = head_dim rounded up to 16; the TMA box is HDP = 64 / 32
// columns wide (one 128- / 64-byte swizzle span per tile row), columns past hds are zero-filled too.
//
// NSTG = 2 or 3 units in flight per SM ("stages": each has its tensor-memory region, its TMA producer warp, its MMA
// issuer warp and its softmax warpgroup).  Two stages: 12 warps --
//   warps 0,3  TMA producers, one per stage (warp 0 also allocates the tensor memory): a Q/K ring and a V
//              ring, so the next unit's Q,K are in flight while this unit's softmax and PV run
//   warps 1,2  MMA issuers, one per stage:   S = QK^T, O = PV
//   warps 4-7  softmax group 0 (units 0, 2, 4, ... of this CTA), TMEM stage 0
//   warps 8-11 softmax group 1 (units 1, 3, 5, ...), TMEM stage 1   (setmaxnreg: 216 registers each,
//              taken from the control warpgroup)
// Three stages (round 2, opt-in PAFUSE_ATT_STAGES=3; the aliased layout of the narrow heads needs 128 + 32 columns per
// stage, so three fit): 20 warps -- producers 0-2, issuers 3-5, softmax groups at warps 8-11 / 12-15 / 16-19 (128
// registers each).  A unit is a latency chain (TMA -> QK^T -> softmax -> 24 small PV MMAs -> output) and the softmax
// groups spend a third of their time waiting for the tensor pipe (profiles/r2h_hotlines_face_attention.txt); the third
// unit fills those waits but costs the operand prefetch distance (one ring slot per stage) and measured slower.
//              thread = tile row = TMEM lane: scores of the row's group read once into registers, max,
//              exp2, sum, fp16 hi/lo -> TMEM; then O -> registers -> 1/sum -> fp16 hi/lo -> per-warp staging
//              rows in shared memory -> 16-byte global stores of whole head slices of [token, C]
#include "kernels.cuh"

#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>

namespace pafuse {

namespace {

constexpr int TILE_ROWS = 128;
__host__ __device__ constexpr int att_threads(int nstg) { return nstg == 2 ? 384 : 640; }
// TMEM columns, two layouts (AttnTcParams::tm_*):
//   aliased (SEP = false): stage s holds S (fp32, 128 columns) at s*128, overwritten in place by P_hi (64 columns
//     of packed fp16 pairs) and P_lo (next 64); O (fp32, HDP columns) at 256 + s*64.  QK^T of unit i+2 can only
//     be issued after PV of unit i, and the softmax group waits for O of its unit before it starts the next one.
//   separate (SEP = true, when O + S + P fit in 256 columns per stage): S, P_hi, P_lo and O do not overlap, so
//     QK^T of unit i+2 is issued as soon as the softmax group has READ S of unit i, and the group writes out O of
//     unit i-2 after the softmax of unit i: it never waits for the tensor pipe (the aliased version spent 45 % of
//     its time waiting for PV, profiles/r1e_*).

struct AttnTcParams {
    int num_tiles;
    int L, Lp, G;             // group length, padded length (multiple of 32), groups per tile (G*Lp <= 128)
    int hd, C;
    int temporal;
    int J, F;
    int tiles_per_seq;        // temporal: ceil(J / G)
    int num_seqs;             // spatial: S*F sequences of L tokens
    long long M;              // valid token rows
    float scale_log2e;        // hd^-0.5 * log2(e)
    int stg_pitch;            // output staging: bytes between rows (= 2*hd + 16)
    int stg_warp_bytes;       // 32 staging rows, rounded up to 128 B
    int chunk_bytes;          // 16, or 8 when a head slice is not 16-byte aligned in [token, C]
    int chunks_per_row;       // 2*hd / chunk_bytes
    int rows_per_iter;        // 32 / chunks_per_row: staging rows one copy-out instruction of a warp covers
    // tensor-memory columns: S / P of stage s at tm_s0 + s * tm_stride_s (+ tm_phi_off / tm_plo_off), O at tm_o0 + s * tm_stride_o
    int tm_s0, tm_stride_s, tm_phi_off, tm_plo_off, tm_o0, tm_stride_o;
    int n_acc;                // O accumulators per stage (HDP columns each): the three f16x3 passes of PV go to different ones
    int o_bufs;               // aliased layout: 2 = O double-buffered per stage (HDP columns each, n_acc = 1): the group writes out unit
                              // i-2 while PV of unit i runs instead of waiting for it; 1 = one buffer, wait for PV before the output
    op_t* o_hi;
    op_t* o_lo;
#ifdef PAFUSE_ABLATE
    int ablate;               // energy ablation builds only (tools/energy_ablation.py): 4 = softmax / output warps skip their work, 8 = no MMAs
#endif
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(sbo_bytes >> 4) << 16;           // leading byte offset (not used by these shapes)
    d |= (uint64_t)(sbo_bytes >> 4) << 32;           // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;                     // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3, int32_t c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// the output staging rows are addressed as shared memory explicitly: through generic pointers the compiler emitted
// generic ST / LD, and the copy-out (generic store -> __syncwarp -> generic load -> STG per iteration) took 2300-4100 of
// the 5700-7000 cycles of a unit (per-phase cycle counts of a -DPAFUSE_ATT_TRACE build, profiles/r2z_*)
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// two values -> packed fp16 hi pair and lo pair.  No range clamp: probabilities are in [0,1], and the
// attention output is a convex combination of V rows the GEMM epilogue already saturated to fp16 range.
__device__ __forceinline__ void split_pair2(float2 v, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v.x, v.y);
    const float2 d = ffma2(__half22float2(h), splat2(-1.0f), v);       // v - hi, exact
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// HDP: tile row width in fp16 elements (64 / 32).  LT: compile-time group length (0 = use p.L; NCH_MAX chunks)
template <int HDP, int LT, bool SEP, int NSTG>
__global__ void __launch_bounds__(att_threads(NSTG), 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                    const AttnTcParams p) {
    constexpr int ROWB = HDP * 2;                              // bytes per tile row = swizzle span
    constexpr int TILE_BYTES = TILE_ROWS * ROWB;
    constexpr int STAGE_BYTES = 6 * TILE_BYTES;                // Qh Ql Kh Kl Vh Vl
    constexpr uint32_t LAYOUT = HDP == 64 ? 2u : 4u;
    constexpr uint32_t SBO = 8 * ROWB;
    constexpr int NCH_MAX = LT ? (LT + 31) / 32 : 4;           // 32-column score chunks of one group
    // operand ring slots: unit `it` loads into slot it % NS.  Two tensor-memory stages bound the units in flight, but
    // the 48 KB stages of the narrow heads leave room for four slots, so Q, K, V are requested three units ahead
    // instead of one and their DRAM latency no longer sits inside the unit's QK -> softmax -> PV chain.
    // A ring slot must always be reused by the SAME stage (NS a multiple of NSTG): the mbarrier waits are parity waits,
    // and a producer of another stage that is two uses behind the slot's barrier would take the completion of use k-2 for
    // that of use k (observed with NSTG = 3 on 4 shared slots: a fast stage lapped a slow one and overwrote live tiles).
    constexpr int NS = NSTG == 3 ? 3 : (HDP == 32 ? 4 : 2);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t qk_full[NS], qk_empty[NS], v_full[NS], v_empty[NS];
    __shared__ __align__(8) uint64_t s_full[NSTG], s_empty[NSTG], p_full[NSTG];
    __shared__ __align__(8) uint64_t o_full[NSTG * 2], o_empty[NSTG * 2];      // [stage * 2 + O buffer]
    constexpr int ATT_THREADS = att_threads(NSTG);
    constexpr int CTRL_WARPS = NSTG == 2 ? 4 : 8;              // warps before the softmax groups (whole warpgroups)
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler: role code uses the uniform datapath
    const int lane = threadIdx.x & 31;
    const int L = LT ? LT : p.L;
    const int Lp = LT ? NCH_MAX * 32 : p.Lp;
    const int nch = LT ? NCH_MAX : p.Lp / 32;
    const int num_units = p.num_tiles * 8;
    const int n_local = (num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int rows_box = p.G * Lp;                             // tile rows every TMA box writes (96 or 128)
    const int key_steps = ((p.G - 1) * Lp + L + 15) / 16;      // 16-key MMA steps that can hold live keys
    // softmax warps (one per 32 tile rows) that hold at least one token row; the others take no part in the
    // barrier protocol (a warp without work could run a unit ahead and arrive twice in one phase)
    int n_live = 0;
    for (int q = 0; q < 4; ++q) {
        const int g = (q * 32) / Lp;
        n_live += (g < p.G && q * 32 - g * Lp < L) ? 1 : 0;
    }

    pdl_launch_dependents();
    // rows no box ever writes must not hold NaN bit patterns (0 * NaN in the PV product)
    for (int i = threadIdx.x; i < NS * STAGE_BYTES / 16; i += ATT_THREADS)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();

    if (warp == 1 && lane == 0) {
        prefetch_tensormap(&tm_hi);
        prefetch_tensormap(&tm_lo);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&qk_full[s], 1);
            mbar_init(&qk_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < NSTG; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], n_live);
            mbar_init(&p_full[s], n_live);                     // one lane per live warp of the stage's softmax group
            for (int b = 0; b < 2; ++b) {
                mbar_init(&o_full[s * 2 + b], 1);
                mbar_init(&o_empty[s * 2 + b], n_live);
            }
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc<1>(&tmem_base_slot, 512);
        tmem_relinquish<1>();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    pdl_wait();                                                // the planes come from the preceding kernel

    // stage a control warp serves: NSTG == 2: producers 0,3 / issuers 1,2; NSTG == 3: producers 0-2 / issuers 3-5
    const bool is_producer = NSTG == 2 ? (warp == 0 || warp == 3) : warp < 3;
    const bool is_issuer = NSTG == 2 ? (warp == 1 || warp == 2) : (warp >= 3 && warp < 6);
    const int ctrl_stage = NSTG == 2 ? (warp == 0 ? 0 : warp == 3 ? 1 : warp - 1) : warp % 3;
    if (warp < CTRL_WARPS) {
    if (NSTG == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");    // control warpgroup(s) donate registers ...
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (is_producer) {
        // ===================== TMA producers: warp 0 feeds ring stage 0 (units 0, 2, ...), warp 3 stage 1 ==========
        // (one thread for both stages made the Q,K loads of unit i+1 queue behind the wait for PV of unit i-2)
        if (lane == 0) {
            const int stage = ctrl_stage;
            const uint32_t tile_tx = (uint32_t)(rows_box * ROWB);          // zero-filled rows count too
            for (int it = stage; it < n_local; it += NSTG) {
                const int u = (int)blockIdx.x + it * (int)gridDim.x;
                const int slot = it % NS;
                const uint32_t ph = (uint32_t)(it / NS) & 1u;
                uint8_t* st = smem + (size_t)slot * STAGE_BYTES;
                const int tile = u >> 3, head = u & 7;
                int ca, cb;                                    // spatial: (first sequence, -) ; temporal: (j0, s)
                if (!p.temporal) {
                    ca = tile * p.G;
                    cb = 0;
                } else {
                    ca = (tile % p.tiles_per_seq) * p.G;
                    cb = tile / p.tiles_per_seq;
                }
#pragma unroll
                for (int w = 0; w < 3; ++w) {                  // q, k, v
                    const int plane = w * 8 + head;
                    if (w == 0) {
                        mbar_wait(&qk_empty[slot], ph ^ 1);    // QK^T of the unit that last used this slot has retired
                        mbar_arrive_expect_tx(&qk_full[slot], 4 * tile_tx);
                    } else if (w == 2) {
                        mbar_wait(&v_empty[slot], ph ^ 1);     // PV of that unit has retired
                        mbar_arrive_expect_tx(&v_full[slot], 2 * tile_tx);
                    }
                    uint64_t* bar = w == 2 ? &v_full[slot] : &qk_full[slot];
                    if (!p.temporal) {
                        tma_load_4d(st + (2 * w) * TILE_BYTES, &tm_hi, bar, 0, 0, ca, plane);
                        tma_load_4d(st + (2 * w + 1) * TILE_BYTES, &tm_lo, bar, 0, 0, ca, plane);
                    } else {
                        tma_load_5d(st + (2 * w) * TILE_BYTES, &tm_hi, bar, 0, 0, ca, cb, plane);
                        tma_load_5d(st + (2 * w + 1) * TILE_BYTES, &tm_lo, bar, 0, 0, ca, cb, plane);
                    }
                }
            }
        }
    } else if (is_issuer) {
        // ===================== MMA issuers, one per stage =====================
        // Per stage the order QK(i), PV(i), QK(i+2), PV(i+2), ... is what lets S alias P; the two stages are
        // independent, and with one issuing thread PV(i+1) used to wait behind the operands of QK(i+2).
        if (lane == 0) {
            const int stage = ctrl_stage;
            const uint32_t idesc_qk = make_idesc_f16(128, (uint32_t)(key_steps * 16));
            const uint32_t idesc_pv = make_idesc_f16(128, HDP) | (1u << 16);      // B (= V) is MN-major
            const uint32_t sa = smem_u32(smem);                // ring slot 0; slot_off() below selects the unit's slot
            const uint32_t d_s = tmem_base + (uint32_t)(p.tm_s0 + stage * p.tm_stride_s);
            const uint32_t d_o = tmem_base + (uint32_t)(p.tm_o0 + stage * p.tm_stride_o);
            // n_acc = 3: the passes go to accumulators 0, 1, 2; n_acc = 2: lo*hi and hi*hi -> 0, hi*lo -> 1; n_acc = 1: all -> 0
            const uint32_t d_o1 = p.n_acc >= 2 ? d_o + (uint32_t)HDP : d_o;
            const uint32_t d_o2 = p.n_acc >= 3 ? d_o + 2u * (uint32_t)HDP : d_o;
            const uint32_t d_phi = d_s + (uint32_t)p.tm_phi_off, d_plo = d_s + (uint32_t)p.tm_plo_off;
            // Operand descriptors of this stage are loop invariants; only their 16-byte address field moves.  Built
            // inside the loops they made the single issuing thread the bottleneck of every attention variant:
            // ~90-110 cycles per issued MMA whatever its shape (profiles/r1m_bench_launches.txt, profiles/r1h_hotlines_attention_sep.txt).
            const uint64_t qh0 = make_desc(sa, SBO, LAYOUT), ql0 = make_desc(sa + TILE_BYTES, SBO, LAYOUT);
            const uint64_t kh0 = make_desc(sa + 2 * TILE_BYTES, SBO, LAYOUT), kl0 = make_desc(sa + 3 * TILE_BYTES, SBO, LAYOUT);
            const uint64_t vh0 = make_desc(sa + 4 * TILE_BYTES, SBO, LAYOUT), vl0 = make_desc(sa + 5 * TILE_BYTES, SBO, LAYOUT);
            constexpr uint64_t V_STEP = (16u * ROWB) >> 4;
            // unit `it` sits in ring slot it % NS: descriptor offset of that slot
            auto slot_off = [&](int it) { return (uint64_t)((it % NS) * (STAGE_BYTES >> 4)); };
            auto issue_qk = [&](int it) {
                const uint32_t ph = (uint32_t)(it / NSTG) & 1u;
                const int slot = it % NS;
                const uint64_t so = slot_off(it);
                mbar_wait(&qk_full[slot], (uint32_t)(it / NS) & 1u);
                if (SEP) mbar_wait(&s_empty[stage], ph ^ 1);   // the softmax group has read S of unit it-2
                tcgen05_fence_after();
                // aliased layout: S[stage] overwrites P[stage] of unit it-2; that PV was issued by this thread
                // before this point and the tensor pipe executes in issue order
#pragma unroll
                for (int k = 0; k < HDP / 16; ++k) {
#ifdef PAFUSE_ABLATE
                    if (p.ablate & 8) break;
#endif
                    const uint64_t ko = (uint64_t)(k * 2);     // 32 bytes along the row, in 16-byte descriptor units
                    umma_f16_ss<1>(d_s, ql0 + so + ko, kh0 + so + ko, idesc_qk, k != 0 ? 1u : 0u);
                    umma_f16_ss<1>(d_s, qh0 + so + ko, kl0 + so + ko, idesc_qk, 1u);
                    umma_f16_ss<1>(d_s, qh0 + so + ko, kh0 + so + ko, idesc_qk, 1u);
                }
                umma_commit<1>(&qk_empty[slot]);               // Q,K tiles of this slot may be overwritten
                umma_commit<1>(&s_full[stage]);
            };
            if (stage < n_local) issue_qk(stage);
            for (int it = stage; it < n_local; it += NSTG) {
                const uint32_t ph = (uint32_t)(it / NSTG) & 1u;
                if (SEP && it + NSTG < n_local) issue_qk(it + NSTG);
                const int slot = it % NS;
                mbar_wait(&v_full[slot], (uint32_t)(it / NS) & 1u);
                // O buffer of this unit and its use count: with two buffers per stage, uses alternate between them
                const int n_use = it / NSTG, ob = p.o_bufs == 2 ? n_use & 1 : 0;
                const uint32_t ph_o = (uint32_t)(p.o_bufs == 2 ? n_use >> 1 : n_use) & 1u;
                const uint32_t ob_off = (uint32_t)(ob * p.n_acc * HDP);
                mbar_wait(&o_empty[stage * 2 + ob], ph_o ^ 1); // the softmax group has read what PV last wrote into this buffer
                mbar_wait(&p_full[stage], ph);                 // P of this unit is in tensor memory
                tcgen05_fence_after();
                uint64_t vh = vh0 + slot_off(it), vl = vl0 + slot_off(it);
                uint32_t ph_a = d_phi, pl_a = d_plo;
                for (int k = 0; k < key_steps; ++k, vh += V_STEP, vl += V_STEP, ph_a += 8u, pl_a += 8u) {
#ifdef PAFUSE_ABLATE
                    if (p.ablate & 8) break;
#endif
                    // 16 keys further down the V tile / 16 fp16 keys = 8 tensor-memory columns further in P
                    // Separate accumulators per pass: the PV MMAs are tiny (N = HDP) and a chain of 3 * key_steps
                    // dependent accumulations into ONE tile ran at the pipe's latency, ~90 cycles per MMA
                    // (profiles/r1m_bench_launches.txt: every attention variant cost ~90-110 cycles per issued MMA).
                    const uint32_t first = k != 0 ? 1u : 0u;
                    umma_f16_ts(d_o + ob_off, pl_a, vh, idesc_pv, first);
                    umma_f16_ts(d_o1 + ob_off, ph_a, vl, idesc_pv, p.n_acc >= 2 ? first : 1u);
                    umma_f16_ts(d_o2 + ob_off, ph_a, vh, idesc_pv, p.n_acc >= 3 ? first : 1u);
                }
                umma_commit<1>(&v_empty[slot]);
                umma_commit<1>(&o_full[stage * 2 + ob]);
                if (!SEP && it + NSTG < n_local) issue_qk(it + NSTG);
            }
        }
    }
    } else {
        if (NSTG == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");   // ... to the softmax warpgroups
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");             // 8*32*40 + 12*32*128 = 59392 <= 640*96 (the CTA's register allocation)
        // ===================== softmax + output =====================
        const int sw = warp - CTRL_WARPS;
        const int wg = sw >> 2;                                // softmax group = TMEM stage
        const int q = warp & 3;                                // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                           // tile row == TMEM lane
        // all 32 rows of a warp belong to one group (Lp is a multiple of 32)
        const int g = (q * 32) / Lp;                           // group of this warp
        const int row0_in_g = q * 32 - g * Lp;                 // first row of the warp inside its group
        const bool warp_live = g < p.G && row0_in_g < L;       // the warp holds at least one token row
        const int c0 = g * nch;                                // first 32-column score chunk of the group
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t s_addr = tmem_base + lane_sel + (uint32_t)(p.tm_s0 + wg * p.tm_stride_s);
        const uint32_t phi_addr = s_addr + (uint32_t)p.tm_phi_off, plo_addr = s_addr + (uint32_t)p.tm_plo_off;
        const uint32_t o_addr = tmem_base + lane_sel + (uint32_t)(p.tm_o0 + wg * p.tm_stride_o);
        const float sc = p.scale_log2e;
        const uint32_t stg_s = smem_u32(smem + NS * STAGE_BYTES + sw * p.stg_warp_bytes);  // this warp's output staging rows
        const uint32_t my_row_s = stg_s + (uint32_t)(lane * p.stg_pitch);
        const int n_zero_chunks = (key_steps + 1) / 2;         // 32-key chunks the PV product reads
        const int cpr = p.chunks_per_row, rpi = p.rows_per_iter;
        const int cp_row = lane / cpr;                         // copy-out role of this lane
        const int cp_cc = lane - cp_row * cpr;
        const bool cp_active = cp_row < rpi;
#ifdef PAFUSE_ATT_TRACE
        // measurement builds only: cycles per phase of the unit loop, printed by one warp of every softmax group of CTA 0
        long long tr[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tr_last = clock64();
#define ATT_TR(i) { const long long t_now = clock64(); tr[i] += t_now - tr_last; tr_last = t_now; }
#else
#define ATT_TR(i)
#endif

        // keys of the other groups: exact zeros in P.  The separate layout never overwrites them: once is enough.
        auto zero_other_groups = [&]() {
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            for (int c = 0; c < n_zero_chunks; ++c)
                if (c < c0 || c >= c0 + nch) {
                    tmem_st_32x16(phi_addr + (uint32_t)(c * 16), z);
                    tmem_st_32x16(plo_addr + (uint32_t)(c * 16), z);
                }
        };
        if (SEP && warp_live) {
            zero_other_groups();
            tmem_st_wait();
        }

        // ---- S -> P of unit `it`; returns the row sum
        auto softmax_unit = [&](int it) -> float {
            const uint32_t ph = (uint32_t)(it / NSTG) & 1u;
            ATT_TR(9)
            mbar_wait(&s_full[wg], ph);
            tcgen05_fence_after();
            ATT_TR(0)
            float sum = 0.f;
#ifdef PAFUSE_ABLATE
            if (p.ablate & 4) {                                // same barrier protocol, no loads / math / stores
                if (SEP) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[wg]);
                    if (warp_live && it >= NSTG) mbar_wait(&o_full[wg * 2], (uint32_t)((it - NSTG) / NSTG) & 1u);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[wg]);
                return 1.f;
            }
#endif
            uint32_t sv[NCH_MAX * 32];
            if (warp_live) {
                // the scores of this row against the keys of its group (this warp's columns), once, into registers
#pragma unroll
                for (int k = 0; k < NCH_MAX; ++k)
                    if (k < nch) tmem_ld_32x32(s_addr + (uint32_t)((c0 + k) * 32), &sv[k * 32]);
                tmem_ld_wait();
            }
            ATT_TR(1)
            if (SEP) {                                         // S may be overwritten by QK^T of unit it+2
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[wg]);
            }
            if (warp_live) {
                // row maximum over the L live keys (four chains)
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int k = 0; k < NCH_MAX; ++k)
                    if (k < nch) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (k * 32 + i < L) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sv[k * 32 + i]));
                    }
                const float moff = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;
                ATT_TR(2)
                // p = exp2(s*c - m*c), row sum, fp16 hi/lo -- on packed pairs (FFMA2 / FADD2, common.cuh): the softmax
                // groups are bound by instruction issue and latency, not by the FMA pipe
                float2 s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                const float2 sc2 = splat2(sc), nm2 = splat2(-moff);
#pragma unroll
                for (int k = 0; k < NCH_MAX; ++k)
                    if (k < nch) {
                        uint32_t hi16[16], lo16[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int col = k * 32 + 2 * i;                 // compile-time masks for LT != 0
                            float2 pr = make_float2(0.f, 0.f);
                            if (col < L) {
                                pr = ffma2(make_float2(__uint_as_float(sv[k * 32 + 2 * i]), __uint_as_float(sv[k * 32 + 2 * i + 1])),
                                           sc2, nm2);
                                pr.x = fast_exp2(pr.x);
                                pr.y = col + 1 < L ? fast_exp2(pr.y) : 0.f;
                                s2[i & 1] = fadd2(s2[i & 1], pr);
                            }
                            split_pair2(pr, hi16[i], lo16[i]);
                        }
                        if (SEP && k == 0 && it >= NSTG) {
                            // P of this unit overwrites P of unit it-2: its PV must have retired (it ran during the
                            // output of unit it-4 and the exponentials above).  This also keeps p_full from running
                            // two phases ahead of the issuer, which would then wait for ever on a parity.
                            ATT_TR(3)
                            mbar_wait(&o_full[wg * 2], (uint32_t)((it - NSTG) / NSTG) & 1u);
                            tcgen05_fence_after();
                            ATT_TR(4)
                        }
                        tmem_st_32x16(phi_addr + (uint32_t)((c0 + k) * 16), hi16);
                        tmem_st_32x16(plo_addr + (uint32_t)((c0 + k) * 16), lo16);
                    }
                sum = (s2[0].x + s2[0].y) + (s2[1].x + s2[1].y);
                if (!SEP) ATT_TR(3)
                if (!SEP) zero_other_groups();
                tmem_st_wait();
            }
            // rows of a dead warp feed stale bits into rows of O nobody stores: nothing to write for them
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[wg]);
            ATT_TR(5)
            return sum;
        };

        // ---- O of unit `it` / sum -> fp16 hi/lo -> global
        // o_seen: the caller has already waited for this unit's o_full phase (separate layout, inside softmax_unit of the
        // unit two later).  Waiting a second time on the same parity would only be safe while PV of the NEXT unit has
        // not completed yet: once it has, that parity names the phase in progress and the wait never returns.
        auto output_unit = [&](int it, float sum, bool o_seen) {
            const int n_use = it / NSTG, ob = p.o_bufs == 2 ? n_use & 1 : 0;
            const uint32_t ph = (uint32_t)(p.o_bufs == 2 ? n_use >> 1 : n_use) & 1u;
            const uint32_t ob_off = (uint32_t)(ob * p.n_acc * HDP);
            const int u = (int)blockIdx.x + it * (int)gridDim.x;
            const int tile = u >> 3, head = u & 7;
            if (!o_seen) mbar_wait(&o_full[wg * 2 + ob], ph);
            tcgen05_fence_after();
            ATT_TR(6)
#ifdef PAFUSE_ABLATE
            if (p.ablate & 4) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[wg * 2 + ob]);
                return;
            }
#endif
            uint32_t ov[HDP];
            auto load_o = [&](uint32_t addr, uint32_t* dst) {
                tmem_ld_32x32(addr, dst);
                if (HDP == 64) tmem_ld_32x32(addr + 32u, dst + 32);
            };
            if (warp_live) {
                load_o(o_addr + ob_off, ov);
                for (int a = 1; a < p.n_acc; ++a) {            // add the other passes' accumulators
                    uint32_t oa[HDP];
                    load_o(o_addr + ob_off + (uint32_t)(a * HDP), oa);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < HDP / 2; ++i) {
                        const float2 t = fadd2(make_float2(__uint_as_float(ov[2 * i]), __uint_as_float(ov[2 * i + 1])),
                                               make_float2(__uint_as_float(oa[2 * i]), __uint_as_float(oa[2 * i + 1])));
                        ov[2 * i] = __float_as_uint(t.x);
                        ov[2 * i + 1] = __float_as_uint(t.y);
                    }
                }
                tmem_ld_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[wg * 2 + ob]);
            ATT_TR(7)
            if (!warp_live) return;
            // token of staging row i of this warp = tok0 + i * tstride, for i < n_rows (all warp-uniform)
            long long tok0;
            int tstride, n_rows = min(32, L - row0_in_g);
            if (!p.temporal) {
                const int seq = tile * p.G + g;
                tok0 = (long long)seq * L + row0_in_g;
                tstride = 1;
                if (seq >= p.num_seqs) n_rows = 0;
            } else {
                const int s = tile / p.tiles_per_seq;
                const int j = (tile % p.tiles_per_seq) * p.G + g;
                tok0 = ((long long)s * p.F + row0_in_g) * p.J + j;
                tstride = p.J;
                if (j >= p.J) n_rows = 0;
            }
            // O / sum -> fp16 hi/lo, once; each half goes through this warp's staging rows (pitch = row bytes + 16)
            // so that the global stores are runs of whole head slices (a lane-per-row store
            // touched 32 different sectors per instruction and cost 25 % of the kernel)
            float inv;                                         // one MUFU instead of the IEEE division sequence: 1 ulp on a scale factor
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sum));
            const float2 inv2 = splat2(inv);
            uint32_t oh[HDP / 2], ol[HDP / 2];
#pragma unroll
            for (int i = 0; i < HDP / 2; ++i)
                split_pair2(fmul2(make_float2(__uint_as_float(ov[2 * i]), __uint_as_float(ov[2 * i + 1])), inv2), oh[i], ol[i]);
            // copy-out: lane -> (row lane / cpr of the current group of rows_per_iter rows, chunk lane % cpr); the
            // per-chunk index arithmetic of the first version (64-bit multiplies per 8 bytes) cost as many
            // instructions as the softmax itself (profiles/r1h_*)
            const size_t row_bytes = (size_t)tstride * (size_t)(p.C * 2);
            const size_t dst0 = (size_t)tok0 * (size_t)(p.C * 2) + (size_t)(head * p.hd * 2 + cp_cc * p.chunk_bytes) +
                                (size_t)cp_row * row_bytes;
            const size_t dst_step = (size_t)rpi * row_bytes;
            ATT_TR(8)
#pragma unroll
            for (int hl = 0; hl < 2; ++hl) {
                __syncwarp();                                  // the previous copy-out has read the staging rows
#pragma unroll
                for (int i = 0; i < HDP / 4; ++i)
                    if (4 * i < p.hd) {
                        if (hl == 0) sts64(my_row_s + 8u * i, oh[2 * i], oh[2 * i + 1]);
                        else sts64(my_row_s + 8u * i, ol[2 * i], ol[2 * i + 1]);
                    }
                __syncwarp();
                uint8_t* dst = reinterpret_cast<uint8_t*>(hl == 0 ? p.o_hi : p.o_lo) + dst0;
                // all shared-memory reads first, then all global stores (at most 8 groups of rows: a head slice is at most
                // 8 chunks): a rolled loop paid the LDS -> STG latency once per group.  The reads are unconditional on a
                // clamped row (inline-asm loads under a lane predicate compile to a branch each); the stores are predicated.
                constexpr int CP_IT = 8;
                const uint32_t src_col = stg_s + (uint32_t)(cp_cc * p.chunk_bytes);
                if (p.temporal) {
                    // token rows J * C * 2 bytes apart: a burst of eight such stores per lane measured 20 % SLOWER than this
                    // rolled loop (face / hands temporal 738 -> 929 us), which paces them
                    uint32_t src = src_col + (uint32_t)(cp_row * p.stg_pitch);
                    const uint32_t src_step = (uint32_t)(rpi * p.stg_pitch);
                    for (int row = cp_row; row < n_rows; row += rpi) {
                        if (cp_active) {
                            if (p.chunk_bytes == 16) *reinterpret_cast<uint4*>(dst) = lds128u(src);
                            else *reinterpret_cast<uint2*>(dst) = lds64(src);
                        }
                        dst += dst_step;
                        src += src_step;
                    }
                } else if (p.chunk_bytes == 16) {
                    uint4 v[CP_IT];
#pragma unroll
                    for (int c = 0; c < CP_IT; ++c) v[c] = lds128u(src_col + (uint32_t)(min(cp_row + c * rpi, 31) * p.stg_pitch));
#pragma unroll
                    for (int c = 0; c < CP_IT; ++c)
                        if (cp_active && cp_row + c * rpi < n_rows) *reinterpret_cast<uint4*>(dst + (size_t)c * dst_step) = v[c];
                } else {
                    uint2 v[CP_IT];
#pragma unroll
                    for (int c = 0; c < CP_IT; ++c) v[c] = lds64(src_col + (uint32_t)(min(cp_row + c * rpi, 31) * p.stg_pitch));
#pragma unroll
                    for (int c = 0; c < CP_IT; ++c)
                        if (cp_active && cp_row + c * rpi < n_rows) *reinterpret_cast<uint2*>(dst + (size_t)c * dst_step) = v[c];
                }
            }
            ATT_TR(10)
        };

        if (!warp_live) {
            // no rows: nothing to compute, nothing to signal
        } else if (!SEP && p.o_bufs != 2) {
            for (int it = wg; it < n_local; it += NSTG) {
                const float sum = softmax_unit(it);
                output_unit(it, sum, false);
            }
        } else {
            // software pipeline: the output of unit it-2 (its PV ran during the softmax of unit it) follows the
            // softmax of unit it.  Aliased layout with two O buffers: S / P of unit it exist only after PV of unit it-2
            // (same columns), so the group cannot start earlier -- but it writes out unit it-2 WHILE PV of unit it and
            // QK^T of unit it+2 run, instead of idling ~2000 cycles per unit on o_full (profiles/r2z_*)
            float prev_sum = 1.f;
            int prev = -1;
            for (int it = wg; it < n_local; it += NSTG) {
                const float sum = softmax_unit(it);
                if (prev >= 0) output_unit(prev, prev_sum, SEP);   // SEP: softmax_unit(it) waited for o_full of unit it - NSTG = prev
                prev = it;
                prev_sum = sum;
            }
            if (prev >= 0) output_unit(prev, prev_sum, false);
        }
#ifdef PAFUSE_ATT_TRACE
        if (blockIdx.x == 0 && lane == 0 && q == 0 && n_local > 0)
            printf("att_trace HDP=%d L=%d SEP=%d grp=%d units=%d cycles/unit: s_full_wait %lld | S_ld %lld | max %lld | exp_split %lld | "
                   "o_full_wait(prev) %lld | P_st+arrive %lld | o_full_wait %lld | O_ld %lld | convert %lld | copyout %lld | other %lld\n",
                   HDP, L, (int)SEP, wg, (n_local - wg + NSTG - 1) / NSTG, tr[0] * NSTG / n_local, tr[1] * NSTG / n_local, tr[2] * NSTG / n_local,
                   tr[3] * NSTG / n_local, tr[4] * NSTG / n_local, tr[5] * NSTG / n_local, tr[6] * NSTG / n_local, tr[7] * NSTG / n_local,
                   tr[8] * NSTG / n_local, tr[10] * NSTG / n_local, tr[9] * NSTG / n_local);
#endif
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 g_enc = nullptr;
int g_n_acc_cap = 1;     // PAFUSE_ATT_NACC: O accumulators per stage the PV passes are spread over (measured: 1 is fastest, the extra tensor-memory reads cost more than the shorter MMA chains save)
int g_stages = 2;        // PAFUSE_ATT_STAGES=3: three units in flight per SM where the tensor-memory layout allows it (narrow heads, aliased
                         // layout).  Opt-in: measured SLOWER (face temporal 1265 -> 1667 us, hands temporal 543 -> 656, hands spatial 627 -> 672,
                         // profiles/r2j_*): with three stages every stage owns ONE ring slot (slots must not be shared between stages, see the
                         // kernel), so Q, K, V of unit i+3 are requested only when unit i has used them -- the two-stage kernel asks three units
                         // ahead, and that prefetch distance is worth more than the third unit in flight
int g_pipe_o = 1;        // PAFUSE_ATT_PIPE=0: aliased layout without the second O buffer (the group waits for PV before it writes out)
int g_sep_mode = 1;      // PAFUSE_ATT_SEP: 0 aliased layout only, 1 separate when it fits (default), 2 also with one group less per tile (slower: measured)

int att_init() {
    if (g_enc) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PAFUSE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        set_last_error("cuTensorMapEncodeTiled not available from the driver");
        return -2;
    }
    g_enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    if (const char* e = getenv("PAFUSE_ATT_SEP")) g_sep_mode = atoi(e);
    if (const char* e = getenv("PAFUSE_ATT_NACC")) g_n_acc_cap = atoi(e);
    if (const char* e = getenv("PAFUSE_ATT_STAGES")) g_stages = atoi(e);
    if (const char* e = getenv("PAFUSE_ATT_PIPE")) g_pipe_o = atoi(e);
    return 0;
}

// planes [24][rows_cap][hds] fp16; the box is hdp >= hds columns and Lp >= L rows per group (everything past
// the tensor extents is zero-filled).
// Spatial:  (hds, L [stride hds], S*F sequences [stride L*hds], 24), box (hdp, Lp, G, 1).
// Temporal: (hds, F [stride J*hds], J [stride hds], S [stride F*J*hds], 24), box (hdp, Lp, G, 1, 1): the box lands
// in shared memory joint-major, i.e. as G groups of Lp consecutive rows.
int make_plane_map(CUtensorMap* map, const op_t* base, long long rows_cap, int hds, int hdp, bool temporal, int J, int F,
                   int S, int G, int L, int Lp) {
    const CUtensorMapSwizzle sw = hdp == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const cuuint64_t e = 2;                                            // bytes per element
    CUresult r;
    if (!temporal) {
        cuuint64_t dims[4] = {(cuuint64_t)hds, (cuuint64_t)L, (cuuint64_t)((long long)S * F), 24};
        cuuint64_t strides[3] = {hds * e, (cuuint64_t)L * hds * e, (cuuint64_t)rows_cap * hds * e};
        cuuint32_t box[4] = {(cuuint32_t)hdp, (cuuint32_t)Lp, (cuuint32_t)G, 1};
        r = g_enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<op_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[5] = {(cuuint64_t)hds, (cuuint64_t)F, (cuuint64_t)J, (cuuint64_t)S, 24};
        cuuint64_t strides[4] = {(cuuint64_t)J * hds * e, hds * e, (cuuint64_t)F * J * hds * e, (cuuint64_t)rows_cap * hds * e};
        cuuint32_t box[5] = {(cuuint32_t)hdp, (cuuint32_t)Lp, (cuuint32_t)G, 1, 1};
        r = g_enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<op_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(attention planes) failed (%d) rows_cap=%lld hds=%d hdp=%d temporal=%d", (int)r,
                       rows_cap, hds, hdp, (int)temporal);
        return -2;
    }
    return 0;
}

template <int HDP, int LT, bool SEP, int NSTG>
int launch_tc(const CUtensorMap& mh, const CUtensorMap& ml, const AttnTcParams& p, cudaStream_t st, int sms) {
    constexpr int NS = NSTG == 3 ? 3 : (HDP == 32 ? 4 : 2);            // operand ring slots (see the kernel)
    const int SMEM = NS * 6 * TILE_ROWS * HDP * 2 + 4 * NSTG * p.stg_warp_bytes + 1024;
    if (SMEM > 227 * 1024) {
        set_last_error("attention_tc: head_dim %d needs %d bytes of shared memory", p.hd, SMEM);
        return -1;
    }
    auto kern = attention_tc_kernel<HDP, LT, SEP, NSTG>;
    static int configured[MAX_DEVICES] = {0};                          // per template instance and per device
    const int dev = current_device_slot();
    if (configured[dev] < SMEM) {
        PAFUSE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured[dev] = SMEM;
    }
    const long long units = (long long)p.num_tiles * 8;
    int grid = (int)(units < sms ? units : sms);
    // on an SM share the CTAs are placed as pairs (no cluster feature is used), so that the share keeps whole
    // TPCs and the CTA pairs of the GEMMs running next to it on other streams still find two free SMs together
    const int cluster = sms < device_sm_count() && grid % 2 == 0 ? 2 : 1;
    PAFUSE_CUDA_OK(launch_chain(kern, dim3((unsigned)grid), dim3(att_threads(NSTG)), (size_t)SMEM, st, cluster, mh, ml, p));
    PAFUSE_LAUNCH_OK();
    return 0;
}

}  // namespace

int launch_attention_tc(const AttnPlanes& pl, op_t* o_hi, op_t* o_lo, int S, int F, int J, int C, int temporal,
                        cudaStream_t st, int sm_limit) {
    if (S == 0) return 0;
    if (int rc = att_init()) return rc;
    const int num_sms = device_sm_count();
    const int sms = sm_limit > 0 && sm_limit < num_sms ? sm_limit : num_sms;
    const int hd = C / 8, hds = attn_head_store(hd), hdp = hds > 32 ? 64 : 32;
    const int L = temporal ? F : J;
    if (hd > 64 || hd % 4 != 0 || L > 128 || pl.hds != hds || pl.rows_cap != (long long)S * F * J) {
        set_last_error("attention_tc: unsupported shape J=%d F=%d C=%d temporal=%d (plane width %d, rows %lld)", J, F, C,
                       temporal, pl.hds, pl.rows_cap);
        return -1;
    }
    AttnTcParams p;
    p.temporal = temporal ? 1 : 0;
    p.L = L;
    p.Lp = (L + 31) / 32 * 32;
    p.G = 128 / p.Lp;
    // tensor-memory layout: separate S / P / O regions when they fit in 256 columns per stage (if necessary with
    // one group less per tile), else the aliased layout
    bool sep = false;
    if (g_sep_mode > 0) {
        for (int G = p.G; G >= 1 && G >= p.G - (g_sep_mode > 1 ? 1 : 0); --G) {
            const int s_cols = G * p.Lp;
            const int key_steps = ((G - 1) * p.Lp + L + 15) / 16;
            const int p_half = (key_steps + 1) / 2 * 16;
            if (hdp + s_cols + 2 * p_half <= 256) {
                sep = true;
                p.G = G;
                p.n_acc = (256 - s_cols - 2 * p_half) / hdp;
                if (p.n_acc > 3) p.n_acc = 3;
                p.tm_o0 = 0;
                p.tm_stride_o = 256;
                p.tm_s0 = p.n_acc * hdp;
                p.tm_stride_s = 256;
                p.tm_phi_off = s_cols;
                p.tm_plo_off = s_cols + p_half;
                break;
            }
        }
    }
    // Aliased layout with the narrow heads: a stage needs 128 (S / P) + 32 (O) columns, so THREE units fit in the 512
    // columns and run in flight (PAFUSE_ATT_STAGES=2 keeps two).
    const bool three = !sep && hdp == 32 && g_stages >= 3;
    if (!sep) {
        p.tm_s0 = 0;
        p.tm_stride_s = 128;
        p.tm_phi_off = 0;
        p.tm_plo_off = 64;
        if (three) {
            p.n_acc = 1;
            p.tm_o0 = 384;
            p.tm_stride_o = hdp;
        } else {
            p.n_acc = hdp == 32 ? 3 : 2;                           // 256 columns are left for O: 2 stages x n_acc x hdp
            p.tm_o0 = 256;
            p.tm_stride_o = p.n_acc * hdp;
        }
    }
    if (p.n_acc > g_n_acc_cap) p.n_acc = g_n_acc_cap < 1 ? 1 : g_n_acc_cap;
    // aliased layout, two stages: the 256 columns behind the S / P regions hold TWO O buffers per stage of n_acc accumulators
    // each, when they fit (one accumulator for the wide heads, up to two for the narrow ones)
    p.o_bufs = 1;
    if (!sep && !three && g_pipe_o && 2 * 2 * p.n_acc * hdp <= 256) {
        p.o_bufs = 2;
        p.tm_stride_o = 2 * p.n_acc * hdp;
    }
#ifdef PAFUSE_ABLATE
    p.ablate = getenv("PAFUSE_ABLATE") ? atoi(getenv("PAFUSE_ABLATE")) : 0;
#endif
    p.hd = hd;
    p.C = C;
    p.J = J;
    p.F = F;
    p.M = (long long)S * F * J;
    p.num_seqs = S * F;
    if (temporal) {
        p.tiles_per_seq = (J + p.G - 1) / p.G;
        p.num_tiles = S * p.tiles_per_seq;
    } else {
        p.tiles_per_seq = 0;
        p.num_tiles = (p.num_seqs + p.G - 1) / p.G;
    }
    p.scale_log2e = (float)(pow((double)hd, -0.5) * 1.4426950408889634);
    p.o_hi = o_hi;
    p.o_lo = o_lo;
    p.stg_pitch = 2 * hd + 16;
    p.stg_warp_bytes = (32 * p.stg_pitch + 127) / 128 * 128;
    p.chunk_bytes = (2 * hd) % 16 == 0 ? 16 : 8;              // hd % 4 == 0, so a slice is at least 8-byte aligned
    p.chunks_per_row = 2 * hd / p.chunk_bytes;
    p.rows_per_iter = 32 / p.chunks_per_row;
    CUtensorMap mh, ml;
    if (int rc = make_plane_map(&mh, pl.hi, pl.rows_cap, hds, hdp, temporal != 0, J, F, S, p.G, L, p.Lp)) return rc;
    if (int rc = make_plane_map(&ml, pl.lo, pl.rows_cap, hds, hdp, temporal != 0, J, F, S, p.G, L, p.Lp)) return rc;
    // the group lengths of the H3WB parts get compile-time masks; anything else runs the generic instance
#define PAFUSE_ATT_CASE(HDPV, LV)                                                             \
    if (hdp == HDPV && (LV == 0 || L == LV))                                                  \
        return sep ? launch_tc<HDPV, LV, true, 2>(mh, ml, p, st, sms)                          \
                   : (three ? launch_tc<HDPV, LV, false, (HDPV == 32 ? 3 : 2)>(mh, ml, p, st, sms) \
                            : launch_tc<HDPV, LV, false, 2>(mh, ml, p, st, sms));
    PAFUSE_ATT_CASE(64, 24)
    PAFUSE_ATT_CASE(64, 27)
    PAFUSE_ATT_CASE(64, 0)
    PAFUSE_ATT_CASE(32, 68)
    PAFUSE_ATT_CASE(32, 42)
    PAFUSE_ATT_CASE(32, 27)
    PAFUSE_ATT_CASE(32, 0)
#undef PAFUSE_ATT_CASE
    set_last_error("attention_tc: no kernel instance for hdp=%d", hdp);
    return -1;
}

// fp32 qkv [M,3C] -> head planes (unit tests; the production path gets the planes from the qkv GEMM epilogue)
__global__ void qkv_to_planes_kernel(const float* __restrict__ qkv, op_t* __restrict__ hi, op_t* __restrict__ lo,
                                     long long M, long long rows_cap, int C, int hd, int hds) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = M * 24 * hds;
    if (idx >= total) return;
    const int d = (int)(idx % hds);
    long long t = idx / hds;
    const long long m = t % M;
    const int plane = (int)(t / M);
    float v = 0.f;
    if (d < hd) v = qkv[(size_t)m * 3 * C + (plane / 8) * C + (plane % 8) * hd + d];
    op_t h, l;
    split_op(v, h, l);
    const size_t o = ((size_t)plane * rows_cap + m) * hds + d;
    hi[o] = h;
    lo[o] = l;
}

int launch_qkv_to_planes(const float* qkv, const AttnPlanes& pl, long long M, int C, cudaStream_t st) {
    const int hd = C / 8;
    const long long total = M * 24 * pl.hds;
    if (total == 0) return 0;
    qkv_to_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(qkv, pl.hi, pl.lo, M, pl.rows_cap, C, hd, pl.hds);
    PAFUSE_LAUNCH_OK();
    return 0;
}

}  // namespace pafuse
