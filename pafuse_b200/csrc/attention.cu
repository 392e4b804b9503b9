// Fused short-sequence attention (mixste.py:63-82, comb=False) over the fixed
// [S,F,J,C] layout: softmax(q k^T * hd^-1/2) v per head, fp32 on CUDA cores.
//
//   spatial  (STE): sequence = the J joints of one (s,f)     -> rows contiguous
//   temporal (TTE): sequence = the F frames of one (s,j)     -> rows strided by J
//
// One CTA = one sequence x HPC heads.  K and V of those heads are staged in
// shared memory with coalesced 16-byte loads; thread (head, query row) keeps its
// q row and output row in registers and streams over the keys with an online
// softmax (4 keys per step, one rescale per step), reading K/V as warp-broadcast
// LDS.128.  The output is written as the fp16 hi/lo pair the proj GEMM consumes.
// Only 3.3 % of the path's FLOPs live here.
#include "kernels.cuh"

#include <math.h>

namespace pafuse {

template <int L, int HD, int HPC, bool TEMPORAL>
__global__ void __launch_bounds__(L* HPC) attention_kernel(AttnParams p) {
    constexpr int HDP = (HD % 32 == 0) ? HD + 4 : HD;       // pad so two heads never share banks
    constexpr int ROW = HPC * HDP;                           // smem floats per key row
    extern __shared__ __align__(16) float smem[];
    float* sk = smem;                                        // [L][ROW]
    float* sv = smem + L * ROW;                              // [L][ROW]

    const int C = p.C;
    const int C3 = 3 * C;
    const int g = blockIdx.x;                                // sequence id
    const int hg = blockIdx.y;                               // head group
    long long row0;
    long long rstride;
    if (TEMPORAL) {
        int s = g / p.J, j = g % p.J;
        row0 = (long long)s * p.F * p.J + j;
        rstride = p.J;
    } else {
        row0 = (long long)g * p.J;
        rstride = 1;
    }
    const float* base = p.qkv + (size_t)row0 * C3;
    const size_t rs = (size_t)rstride * C3;
    const int hoff = hg * HPC * HD;                          // channel offset of this head group

    // ---- stage K and V (coalesced float4)
    constexpr int V4_PER_ROW = HPC * HD / 4;
    constexpr int HD4 = HD / 4;
    for (int idx = threadIdx.x; idx < L * V4_PER_ROW; idx += L * HPC) {
        int r = idx / V4_PER_ROW, w = idx % V4_PER_ROW;
        int hl = w / HD4, d4 = w % HD4;
        const float* src = base + r * rs + hoff + hl * HD + d4 * 4;
        float4 kk = *reinterpret_cast<const float4*>(src + C);
        float4 vv = *reinterpret_cast<const float4*>(src + 2 * C);
        *reinterpret_cast<float4*>(sk + r * ROW + hl * HDP + d4 * 4) = kk;
        *reinterpret_cast<float4*>(sv + r * ROW + hl * HDP + d4 * 4) = vv;
    }

    const int hl = threadIdx.x / L;                          // local head
    const int i = threadIdx.x % L;                           // query row
    float q[HD];
    {
        const float* qs = base + i * rs + hoff + hl * HD;
#pragma unroll
        for (int d4 = 0; d4 < HD4; ++d4) {
            float4 t = *reinterpret_cast<const float4*>(qs + d4 * 4);
            q[d4 * 4 + 0] = t.x;
            q[d4 * 4 + 1] = t.y;
            q[d4 * 4 + 2] = t.z;
            q[d4 * 4 + 3] = t.w;
        }
    }
    __syncthreads();

    const float scale = p.scale;                             // head_dim^-0.5 (mixste.py:51)
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float mrun = -INFINITY, lrun = 0.f;
    const float* kh = sk + hl * HDP;
    const float* vh = sv + hl * HDP;

    constexpr int KB = 4;
    for (int j0 = 0; j0 < L; j0 += KB) {
        float s[KB];
#pragma unroll
        for (int t = 0; t < KB; ++t) s[t] = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < HD4; ++d4) {
#pragma unroll
            for (int t = 0; t < KB; ++t) {
                int j = j0 + t < L ? j0 + t : L - 1;
                float4 kk = *reinterpret_cast<const float4*>(kh + j * ROW + d4 * 4);
                s[t] = fmaf(q[d4 * 4 + 0], kk.x, s[t]);
                s[t] = fmaf(q[d4 * 4 + 1], kk.y, s[t]);
                s[t] = fmaf(q[d4 * 4 + 2], kk.z, s[t]);
                s[t] = fmaf(q[d4 * 4 + 3], kk.w, s[t]);
            }
        }
        float mnew = mrun;
#pragma unroll
        for (int t = 0; t < KB; ++t) {
            s[t] = (j0 + t < L) ? s[t] * scale : -INFINITY;
            mnew = fmaxf(mnew, s[t]);
        }
        float corr = expf(mrun - mnew);                      // 0 on the first step (mrun = -inf)
        float pr[KB];
        float psum = 0.f;
#pragma unroll
        for (int t = 0; t < KB; ++t) {
            pr[t] = expf(s[t] - mnew);                       // exp(-inf) = 0 for masked keys
            psum += pr[t];
        }
        lrun = lrun * corr + psum;
        mrun = mnew;
#pragma unroll
        for (int d4 = 0; d4 < HD4; ++d4) {
            float a0 = o[d4 * 4 + 0] * corr, a1 = o[d4 * 4 + 1] * corr, a2 = o[d4 * 4 + 2] * corr,
                  a3 = o[d4 * 4 + 3] * corr;
#pragma unroll
            for (int t = 0; t < KB; ++t) {
                int j = j0 + t < L ? j0 + t : L - 1;
                float4 vv = *reinterpret_cast<const float4*>(vh + j * ROW + d4 * 4);
                a0 = fmaf(pr[t], vv.x, a0);
                a1 = fmaf(pr[t], vv.y, a1);
                a2 = fmaf(pr[t], vv.z, a2);
                a3 = fmaf(pr[t], vv.w, a3);
            }
            o[d4 * 4 + 0] = a0;
            o[d4 * 4 + 1] = a1;
            o[d4 * 4 + 2] = a2;
            o[d4 * 4 + 3] = a3;
        }
    }
    const float inv = 1.0f / lrun;
    const size_t orow = (size_t)(row0 + (long long)i * rstride) * C + hoff + hl * HD;
#pragma unroll
    for (int d4 = 0; d4 < HD4; ++d4) {
        float v4[4] = {o[d4 * 4 + 0] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv};
        uint2 hi, lo;
        split4(v4, hi, lo);
        *reinterpret_cast<uint2*>(p.out_hi + orow + d4 * 4) = hi;
        *reinterpret_cast<uint2*>(p.out_lo + orow + d4 * 4) = lo;
    }
}

template <int L, int HD, int HPC, bool TEMPORAL>
static int launch_one(const AttnParams& p, cudaStream_t st) {
    constexpr int HDP = (HD % 32 == 0) ? HD + 4 : HD;
    size_t smem = (size_t)2 * L * HPC * HDP * sizeof(float);
    auto kern = attention_kernel<L, HD, HPC, TEMPORAL>;
    static bool configured[MAX_DEVICES] = {false};           // per template instance and per device
    const int dev = current_device_slot();
    if (!configured[dev]) {
        PAFUSE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    long long groups = TEMPORAL ? (long long)p.S * p.J : (long long)p.S * p.F;
    dim3 grid((unsigned)groups, 8 / HPC);
    kern<<<grid, L * HPC, smem, st>>>(p);
    PAFUSE_LAUNCH_OK();
    return 0;
}

int launch_attention(const AttnParams& p_in, cudaStream_t st) {
    if (p_in.S == 0) return 0;
    AttnParams p = p_in;
    const int hd = p.C / 8;
    p.scale = (float)pow((double)hd, -0.5);
    // shapes of the three PAFUSE part denoisers (diffusionpose.py:141, h3wb_dataset.py:198-213), F = 27
    if (!p.temporal) {
        if (p.J == 24 && hd == 48) return launch_one<24, 48, 4, false>(p, st);
        if (p.J == 68 && hd == 28) return launch_one<68, 28, 2, false>(p, st);
        if (p.J == 42 && hd == 32) return launch_one<42, 32, 4, false>(p, st);
    } else if (p.F == 27) {
        if (hd == 48) return launch_one<27, 48, 4, true>(p, st);
        if (hd == 28) return launch_one<27, 28, 4, true>(p, st);
        if (hd == 32) return launch_one<27, 32, 4, true>(p, st);
    }
    set_last_error("attention: unsupported shape J=%d F=%d C=%d temporal=%d", p.J, p.F, p.C, p.temporal);
    return -1;
}

}  // namespace pafuse
