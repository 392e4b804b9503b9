// Shared device helpers for the sm_100a kernels: error plumbing, fp16 hi/lo
// splitting, and thin inline-PTX wrappers for mbarrier / TMA / tcgen05.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pafuse {

// ---------------------------------------------------------------- errors
void set_last_error(const char* fmt, ...);
extern thread_local long long g_launch_count;  // kernels launched by this library on this thread

#define PAFUSE_CUDA_OK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::pafuse::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                                     cudaGetErrorString(_e));                             \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

#define PAFUSE_LAUNCH_OK()                                                                \
    do {                                                                                  \
        ++::pafuse::g_launch_count;                                                       \
        cudaError_t _e = cudaPeekAtLastError();                                           \
        if (_e != cudaSuccess) {                                                          \
            ::pafuse::set_last_error("%s:%d: launch failed -> %s", __FILE__, __LINE__,    \
                                     cudaGetErrorString(_e));                             \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

// ---------------------------------------------------------------- per-device state
// Function attributes (the dynamic shared-memory opt-in) and the SM count belong to a DEVICE, not to the process:
// a context may be created on cuda:1 after cuda:0 was used (one process driving several GPUs, the reference's
// nn.DataParallel mode).  Everything that was a process-wide "configured" flag is indexed by the current device.
constexpr int MAX_DEVICES = 64;
inline int current_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}
inline int device_sm_count() {
    static int sms[MAX_DEVICES] = {0};
    const int d = current_device_slot();
    if (sms[d] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 1;
        sms[d] = v;
    }
    return sms[d];
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the denoiser chain is launched with programmaticStreamSerialization and runs
//   pdl_launch_dependents();  ... set-up that touches no global data ...  pdl_wait();
// so the next kernel's CTAs become resident, and run their prologue (barrier init, tensor-memory allocation,
// tensor-map prefetch, shared-memory clearing), while the tail of this kernel drains.  pdl_wait() returns once
// the preceding grid has completed and its writes are visible; without the launch attribute it is a no-op.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();                             // PAFUSE_PDL=0 turns the launch attribute off

// <<<grid, block, smem, st>>> with the programmatic-serialization attribute (and an optional cluster size)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                                Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = (unsigned)na;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- fp16 hi/lo split
// Tensor-core operand format of the whole path ("f16x3"): an fp32 value v is carried as
//   hi = fp16(v), lo = fp16(v - hi)      ->  hi + lo keeps ~22 significant bits,
// and a product a*w is evaluated as a_lo*w_hi + a_hi*w_lo + a_hi*w_hi (3 tensor-core passes,
// fp32 accumulation); the dropped lo*lo term is ~2^-22 relative.  fp16 (11-bit significand) is
// used instead of bf16 (8-bit) because both run at the same tcgen05 rate and the pair is 64x
// more precise; its narrow range is handled by (a) saturating the conversion and (b) storing
// GEMM weights pre-multiplied by WEIGHT_SCALE (a power of two, undone exactly in the epilogue)
// so that the lo halves of typical |w| ~ 1e-2 stay out of the fp16 subnormal range.
typedef __half op_t;
constexpr float WEIGHT_SCALE = 256.0f;
constexpr float WEIGHT_UNSCALE = 1.0f / 256.0f;

__device__ __forceinline__ void split_op(float v, op_t& hi, op_t& lo) {
    v = v != v ? v : fminf(fmaxf(v, -65504.0f), 65504.0f);      // saturate, but let a NaN stay a NaN like the fp32 reference
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
// torch.clamp semantics: a NaN stays a NaN (fminf / fmaxf alone would return the bound)
__device__ __forceinline__ float clamp_keep_nan(float v, float lo, float hi) { return v != v ? v : fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float join_op(op_t hi, op_t lo) { return __half2float(hi) + __half2float(lo); }

__device__ __forceinline__ uint32_t pack_op2(op_t a, op_t b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// split 4 floats, produce 2x uint2 (hi, lo) of packed fp16
__device__ __forceinline__ void split4(const float v[4], uint2& hi, uint2& lo) {
    op_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_op(v[i], h[i], l[i]);
    hi.x = pack_op2(h[0], h[1]);
    hi.y = pack_op2(h[2], h[3]);
    lo.x = pack_op2(l[0], l[1]);
    lo.y = pack_op2(l[2], l[3]);
}

// two values -> packed fp16 hi pair and lo pair (a in the low halves), conversions saturating to the fp16
// range: 6 instructions per pair (F2FP.SATFINITE.PACK, 2 unpacks, 2 FADD, F2FP) instead of 14 for two split_op
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void split_pair_sat(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = cvt_f16x2_sat(a, b);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = cvt_f16x2_sat(a - hf.x, b - hf.y);
}

// ---------------------------------------------------------------- packed fp32 pairs (FFMA2 / FMUL2 / FADD2)
// sm_100 executes two fp32 operations per lane in one instruction on 64-bit register pairs.  The GEMM epilogues are
// bound by instruction ISSUE (IPC ~0.5 per scheduler on ~21 instructions per element, profiles/r2d_*), not by the FMA
// pipe, so halving the arithmetic instruction count is what shortens them.  Every operation is the IEEE operation of
// its scalar counterpart on each half: results are bit-identical to the scalar formulation.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
    return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// split_pair_sat with the subtraction as one packed instruction: 5 instructions per pair
__device__ __forceinline__ void split_pair_sat2(float2 v, uint32_t& hi, uint32_t& lo) {
    hi = cvt_f16x2_sat(v.x, v.y);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 d = ffma2(hf, splat2(-1.0f), v);                     // v - hf, exact (the difference is representable)
    lo = cvt_f16x2_sat(d.x, d.y);
}

__device__ __forceinline__ float gelu_erf(float x) {  // nn.GELU() default (exact erf form)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// The same function in 15 instructions for the GEMM epilogue (erff + the 0.5x(1+.) form cost ~30 and made the
// fc1 epilogue slower than its main loop): with z = |x|/sqrt(2), erfc(z) = poly5(t) exp(-z^2), t = 1/(1 + p z)
// (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7), and gelu(x) = max(x,0) - 0.5 |x| erfc(z), which has no
// cancellation on either side.  Max abs deviation from the exact function 3.3e-7 over [-12,12] (the fp32
// evaluation of the textbook form deviates by 4.5e-7).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float ax = fabsf(x);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.3275911f * 0.70710678118654752440f, 1.0f)));
    float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    p *= t;                                                          // 0.5 * poly5(t)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * -0.72134752044448170368f) * x));   // exp(-x^2/2)
    return fmaxf(x, 0.0f) - (ax * p) * e;
}

// gelu_erf_fast on a pair: 18 instructions per pair (2 abs, 8 packed FMA / MUL, 4 MUFU, 2 max, ...) instead of 30.
// Same polynomial; the last step is one fused multiply-add (max(x,0) - (|x| p) e), i.e. one rounding less.
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    float2 t = ffma2(ax, splat2(0.3275911f * 0.70710678118654752440f), splat2(1.0f));
    asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(t.x));
    asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(t.y));
    float2 p = ffma2(t, splat2(-0.5f * 1.061405429f), splat2(-0.5f * -1.453152027f));   // -0.5 * poly5(t): the sign is folded in
    p = ffma2(p, t, splat2(-0.5f * 1.421413741f));
    p = ffma2(p, t, splat2(-0.5f * -0.284496736f));
    p = ffma2(p, t, splat2(-0.5f * 0.254829592f));
    p = fmul2(p, t);
    float2 e = fmul2(fmul2(x, splat2(-0.72134752044448170368f)), x);                     // -x^2/2 * log2(e)
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e.x));
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e.y));
    return ffma2(fmul2(ax, p), e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- PTX: misc
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 %%rx;\n\t"
        ".reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %1;\n\t"
        "@%%px mov.s32 %0, 1;\n\t"
        "}\n"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// ---------------------------------------------------------------- PTX: mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a wrong barrier protocol must turn into a trap (reported as a
// launch error), never into a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) {
            printf("pafuse: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            asm volatile("trap;");
        }
    }
}

// ---------------------------------------------------------------- PTX: TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load global -> shared, completion on an mbarrier (tx bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// TMA stores (shared -> global) of one box, tracked by bulk async-groups of the issuing thread.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// global[box] += shared[box] (fp32 add performed at the L2): the residual update x += y without reading x into the SM
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {   // <= N groups may still be reading their shared source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {        // <= N groups not yet complete (writes performed)
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- PTX: cluster
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of this cluster.  Relaxed: the arrivals of this
// library only hand tensor-memory buffers back to the MMA issuer and are ordered by tcgen05.fence; with
// .release.cluster the compiler emits ERRBAR + CGAERRBAR in front of every arrive, which accounted for a
// quarter of the GEMM epilogue warps' time (profiles/r1e_*).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's smem, the transaction bytes are counted on the
// mbarrier of the pair's even CTA (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------- PTX: tcgen05 (CG = cta_group, 1 or 2)
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    if (CG == 1)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if (CG == 1)
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], 16-bit inputs, fp32 accumulate, issued by ONE thread
// (of the pair's even CTA when CG == 2: the instruction then reads both CTAs' shared memory at
// the same offsets and writes both CTAs' tensor memory).
template <int CG>
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// arrive on an mbarrier (same smem offset in every CTA of the group) once all previously issued
// MMAs of this thread have completed
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address, 16-byte units
    d |= (uint64_t)0 << 16;                          // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024u >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // layout: SWIZZLE_128B
    return d;
}

// K-major swizzled operand descriptor with an explicit swizzle span: layout 2 = 128 B, 4 = 64 B, 6 = 32 B;
// sbo_bytes = distance between 8-row groups (8 x the row pitch)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major.
// M is the instruction's M (128 for cta_group::1, 256 for cta_group::2).
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)          // D format: f32
           | (0u << 7)        // A format: f16
           | (0u << 10)       // B format: f16
           | ((N >> 3) << 17) // N, 3 LSBs dropped
           | ((M >> 4) << 24);// M, 4 LSBs dropped
}

}  // namespace pafuse
