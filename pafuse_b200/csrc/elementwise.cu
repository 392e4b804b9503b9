// HBM-bound kernels of the denoising path: embedding, time-MLP, LayerNorm(+split),
// head, DDIM x0/eps update, part re-assembly and multi-hypothesis aggregation.
//
// Activation layout everywhere: token row m = (s*F + f)*J + j of a [S,F,J,C] tensor
// (s = sequence = clip x hypothesis x {orig,flip}); no transposes between spatial
// and temporal blocks (replaces the 16 rearrange copies of mixste.py:244-274).
#include "kernels.cuh"

namespace pafuse {

// ------------------------------------------------------------------ weight split
__global__ void split_weights_kernel(const float* __restrict__ w, op_t* __restrict__ hi,
                                     op_t* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        op_t h, l;
        split_op(w[i] * WEIGHT_SCALE, h, l);           // exact power-of-two pre-scale, undone in the GEMM epilogue
        hi[i] = h;
        lo[i] = l;
    }
}

int launch_split_weights(const float* w, op_t* hi, op_t* lo, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    split_weights_kernel<<<blocks, 256, 0, st>>>(w, hi, lo, n);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// qkv weight [3C,C] (rows [q|k|v], head-major, mixste.py:65) -> [24*hds, C]: row (plane, d) = source row
// plane*hd + d for d < hd, zero otherwise (plane = which*8 + head).  With zero weight rows and zero bias the
// pad columns of the GEMM output are exact zeros, which is what the attention tiles need.
__global__ void pack_qkv_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ wp,
                                float* __restrict__ bp, int C, int hd, int hds) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)24 * hds * C;
    if (i >= total) return;
    int k = (int)(i % C);
    int row = (int)(i / C);
    int plane = row / hds, d = row % hds;
    wp[i] = d < hd ? w[(size_t)(plane * hd + d) * C + k] : 0.f;
    if (k == 0) bp[row] = d < hd ? b[plane * hd + d] : 0.f;
}

int launch_pack_qkv(const float* w, const float* b, float* wp, float* bp, int C, int hd, int hds, cudaStream_t st) {
    size_t total = (size_t)24 * hds * C;
    pack_qkv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, b, wp, bp, C, hd, hds);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ time MLP
// temb = W2 * gelu(W1 * sinus + b1) + b2      (mixste.py:179-184); one CTA per part-call.
__global__ void time_mlp_kernel(const float* __restrict__ sinus, const float* __restrict__ w1,
                                const float* __restrict__ b1, const float* __restrict__ w2,
                                const float* __restrict__ b2, float* __restrict__ temb, int C) {
    extern __shared__ float sm[];
    float* e = sm;           // [C]
    float* h = sm + C;       // [2C]
    for (int i = threadIdx.x; i < C; i += blockDim.x) e[i] = sinus[i];
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < 2 * C; o += nw) {
        const float* wr = w1 + (size_t)o * C;
        float acc = 0.f;
        for (int i = lane; i < C; i += 32) acc = fmaf(wr[i], e[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) h[o] = gelu_erf(acc + b1[o]);
    }
    __syncthreads();
    for (int o = warp; o < C; o += nw) {
        const float* wr = w2 + (size_t)o * 2 * C;
        float acc = 0.f;
        for (int i = lane; i < 2 * C; i += 32) acc = fmaf(wr[i], h[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) temb[o] = acc + b2[o];
    }
}

int launch_time_mlp(const float* sinus, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* temb, int C, cudaStream_t st) {
    time_mlp_kernel<<<1, 512, 3 * C * sizeof(float), st>>>(sinus, w1, b1, w2, b2, temb, C);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ embedding
// x[m,:] = We * [u,v,x,y,z] + be + Spatial_pos_embed[j] + temb    (mixste.py:227-235)
// The 3D input is taken from the sampler state `img` (clamped/scaled when
// apply_clamp, diffusionpose.py:193-194) and, for flip sequences (s >= R_flip_start),
// x is negated and left/right joints are swapped on the fly (:195-198).
// One warp per token row; lanes stride over channels.
template <int NV>
__device__ __forceinline__ void ln_row_stats(const float4 (&v)[NV], int lane, int C, float& mean, float& rstd, float eps) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);   // lanes past C hold zeros
    const float invC = 1.0f / (float)C;
    mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (4 * lane + 128 * i < C) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    rstd = 1.0f / sqrtf(warp_sum(q) * invC + eps);
}

// One warp per token row; the row lives in registers as NV float4 per lane (channels 4*lane + 128*i, the layout of
// ln_chain_kernel): 16-byte loads of the 20 weights of four channels and of the bias / positional / time vectors (the
// round-1 kernel did eight scalar loads per element and wrote x at 2 TB/s), one 16-byte store per four channels.  When
// g1 is given, norm1 of STE block 0 (mixste.py:114) is applied to the row while it is in registers and its fp16 hi/lo
// pair is written as well: the separate ln_chain launch of block 0 (a read of x and a launch per part) is gone.  The
// arithmetic per element and the statistics are those of the two separate kernels, bit for bit.
template <int NV>
__global__ void __launch_bounds__(256) embed_kernel(EmbedParams p) {
    pdl_launch_dependents();
    pdl_wait();
    // persistent warps: with one row per warp and 8 rows per block the launch was bound by block scheduling
    // (147 K blocks of ~2 us on the face part), not by memory
    const int warps_per_block = blockDim.x >> 5;
    const long long warp_stride = (long long)gridDim.x * warps_per_block;
    const int lane = threadIdx.x & 31;
    for (long long m = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); m < p.M; m += warp_stride) {
    int j = (int)(m % p.J);
    long long sf = m / p.J;
    int f = (int)(sf % p.F);
    int s = (int)(sf / p.F) + p.s0;               // global sequence id
    bool flip = s >= p.R;                          // second half of the sequence set = flip-TTA twins
    int r = flip ? s - p.R : s;                    // (clip, hypothesis) id
    int b = r / p.H;
    int g = p.part_joints[j];                      // whole-body joint id
    int gsrc = flip ? p.flip_perm[g] : g;
    const float* x2 = (flip ? p.x2d_flip : p.x2d) + (((size_t)b * p.F + f) * p.num_kps + g) * 2;
    const float* x3 = p.x3d + (((size_t)r * p.F + f) * p.num_kps + gsrc) * 3;
    float in[5];
    in[0] = x2[0];
    in[1] = x2[1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = x3[c];
        if (p.apply_clamp) {
            v = clamp_keep_nan(v, -p.clamp, p.clamp);
            v = __fdiv_rn(v, p.scale);
        }
        in[2 + c] = v;
    }
    if (flip) in[2] = -in[2];
    const int C = p.C;
    float* xo = p.x + (size_t)m * C;
    const float* pos = p.spos + (size_t)j * C;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 4 * lane + 128 * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) {
            float w[20];                                             // We[c .. c+3][0 .. 4], contiguous
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.we + (size_t)c * 5) + q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
            const float4 be = __ldg(reinterpret_cast<const float4*>(p.be + c));
            const float4 po = __ldg(reinterpret_cast<const float4*>(pos + c));
            const float4 te = __ldg(reinterpret_cast<const float4*>(p.temb + c));
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < 5; ++q) acc = fmaf(in[q], w[5 * e + q], acc);
                o[e] = acc;
            }
            v[i].x = ((o[0] + be.x) + po.x) + te.x;
            v[i].y = ((o[1] + be.y) + po.y) + te.y;
            v[i].z = ((o[2] + be.z) + po.z) + te.z;
            v[i].w = ((o[3] + be.w) + po.w) + te.w;
            *reinterpret_cast<float4*>(xo + c) = v[i];
        }
    }
    if (p.g1) {
        float mean, rstd;
        ln_row_stats<NV>(v, lane, C, mean, rstd, p.eps1);
        op_t* oh = p.out_hi + (size_t)m * C;
        op_t* ol = p.out_lo + (size_t)m * C;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 4 * lane + 128 * i;
            if (c < C) {
                const float4 gg = __ldg(reinterpret_cast<const float4*>(p.g1 + c));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + c));
                float y[4];
                y[0] = (v[i].x - mean) * rstd * gg.x + bb.x;
                y[1] = (v[i].y - mean) * rstd * gg.y + bb.y;
                y[2] = (v[i].z - mean) * rstd * gg.z + bb.z;
                y[3] = (v[i].w - mean) * rstd * gg.w + bb.w;
                uint2 h, l;
                split4(y, h, l);
                *reinterpret_cast<uint2*>(oh + c) = h;
                *reinterpret_cast<uint2*>(ol + c) = l;
            }
        }
    }
    }
}

// rows-per-warp kernels: exactly as many blocks as are resident at once (occupancy x SMs), each warp striding over rows
// -- a larger grid would run its last blocks after the first ones have finished their whole stride loop
template <auto KERN>
static unsigned persistent_blocks(long long rows, int wpb) {
    auto kern = KERN;
    static int cached[MAX_DEVICES] = {0};                            // per kernel instance (non-type template argument) and device
    int& per_sm = cached[current_device_slot()];
    if (per_sm == 0 &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, 0) != cudaSuccess || per_sm < 1))
        per_sm = 4;
    const long long want = (rows + wpb - 1) / wpb, cap = (long long)device_sm_count() * per_sm;
    return (unsigned)(want < cap ? want : cap);
}

int launch_embed(const EmbedParams& p, cudaStream_t st) {
    if (p.M == 0) return 0;
    if (p.C % 4 != 0 || p.C > 384) {
        set_last_error("embed: C=%d unsupported (multiple of 4, <= 384)", p.C);
        return -1;
    }
    const int wpb = 8;
    if (p.C <= 256)
        PAFUSE_CUDA_OK(launch_chain(embed_kernel<2>, dim3(persistent_blocks<embed_kernel<2>>(p.M, wpb)), dim3(wpb * 32), 0, st, 1, p));
    else
        PAFUSE_CUDA_OK(launch_chain(embed_kernel<3>, dim3(persistent_blocks<embed_kernel<3>>(p.M, wpb)), dim3(wpb * 32), 0, st, 1, p));
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ LayerNorm (+ chained norm) + fp16 split
// Optional first stage (g0 != nullptr):  x <- LN(x; g0,b0,eps0) [+ add_f[f,:]]   written back (residual stream)
//     = the shared Spatial_norm / Temporal_norm after every block (mixste.py:243,257,269,273)
//       and the Temporal_pos_embed add before TTE block 0 (:250).
// Second stage (g1 != nullptr):         a  = LN(x; g1,b1,eps1) -> fp16 hi/lo     (norm1 / norm2 of the next GEMM)
// One warp per row, row kept in registers as NV float4 per lane (channels 4*lane + 128*i): 16-byte loads
// and stores of x, 8-byte stores of the fp16 halves.
template <int NV>
__global__ void __launch_bounds__(256) ln_chain_kernel(LnParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int warps_per_block = blockDim.x >> 5;
    const long long m = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (m >= p.M) return;
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    float* xr = p.x + (size_t)m * C;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 4 * lane + 128 * i;
        v[i] = c < C ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p.g0) {
        float mean, rstd;
        ln_row_stats<NV>(v, lane, C, mean, rstd, p.eps0);
        const float* addr = nullptr;
        if (p.add_f) {
            const int f = (int)((m / p.J) % p.F);
            addr = p.add_f + (size_t)f * C;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 4 * lane + 128 * i;
            if (c < C) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(p.g0 + c));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.b0 + c));
                float4 y;
                y.x = (v[i].x - mean) * rstd * g.x + b.x;
                y.y = (v[i].y - mean) * rstd * g.y + b.y;
                y.z = (v[i].z - mean) * rstd * g.z + b.z;
                y.w = (v[i].w - mean) * rstd * g.w + b.w;
                if (addr) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(addr + c));
                    y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w;
                }
                v[i] = y;
                *reinterpret_cast<float4*>(xr + c) = y;
            }
        }
    }
    if (p.g1) {
        float mean, rstd;
        ln_row_stats<NV>(v, lane, C, mean, rstd, p.eps1);
        op_t* oh = p.out_hi + (size_t)m * C;
        op_t* ol = p.out_lo + (size_t)m * C;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 4 * lane + 128 * i;
            if (c < C) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(p.g1 + c));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.b1 + c));
                float y[4];
                y[0] = (v[i].x - mean) * rstd * g.x + b.x;
                y[1] = (v[i].y - mean) * rstd * g.y + b.y;
                y[2] = (v[i].z - mean) * rstd * g.z + b.z;
                y[3] = (v[i].w - mean) * rstd * g.w + b.w;
                uint2 h, l;
                split4(y, h, l);
                *reinterpret_cast<uint2*>(oh + c) = h;
                *reinterpret_cast<uint2*>(ol + c) = l;
            }
        }
    }
}

int launch_ln_chain(const LnParams& p, cudaStream_t st) {
    if (p.M == 0) return 0;
    const int wpb = 8;
    unsigned blocks = (unsigned)((p.M + wpb - 1) / wpb);
    if (p.C % 4 != 0 || p.C > 384) {
        set_last_error("ln_chain: C=%d unsupported (multiple of 4, <= 384)", p.C);
        return -1;
    }
    if (p.C <= 256)
        PAFUSE_CUDA_OK(launch_chain(ln_chain_kernel<2>, dim3(blocks), dim3(wpb * 32), 0, st, 1, p));
    else
        PAFUSE_CUDA_OK(launch_chain(ln_chain_kernel<3>, dim3(blocks), dim3(wpb * 32), 0, st, 1, p));
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ head
// y = Linear(C,3)( LN_head( LN_shared(x) ) )      (mixste.py:273, :207-210, :291)
// written straight into the whole-body prediction tensor [S,F,num_kps,3] at the
// part's joint ids (replaces torch.cat, diffusionpose.py:165-171).
// One warp per row, the row in registers as NV float4 per lane (16-byte loads; the round-1 kernel read x with scalar
// loads at 1.8-2.2 TB/s).
template <int NV>
__global__ void __launch_bounds__(256) head_kernel(HeadParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int warps_per_block = blockDim.x >> 5;
    const long long warp_stride = (long long)gridDim.x * warps_per_block;
    const int lane = threadIdx.x & 31;
    const int C = p.C;
    for (long long m = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); m < p.M; m += warp_stride) {
    const float* xr = p.x + (size_t)m * C;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 4 * lane + 128 * i;
        v[i] = c < C ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {
        const float* g = stage == 0 ? p.g0 : p.g1;
        const float* bb = stage == 0 ? p.b0 : p.b1;
        const float eps = stage == 0 ? p.eps0 : p.eps1;
        if (!g) continue;
        float mean, rstd;
        ln_row_stats<NV>(v, lane, C, mean, rstd, eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = 4 * lane + 128 * i;
            if (c < C) {
                const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bb + c));
                v[i].x = (v[i].x - mean) * rstd * gg.x + b4.x;
                v[i].y = (v[i].y - mean) * rstd * gg.y + b4.y;
                v[i].z = (v[i].z - mean) * rstd * gg.z + b4.z;
                v[i].w = (v[i].w - mean) * rstd * gg.w + b4.w;
            }
        }
    }
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = 4 * lane + 128 * i;
        if (c < C) {
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(p.wh + (size_t)o * C + c));
                acc[o] = fmaf(v[i].x, w.x, fmaf(v[i].y, w.y, fmaf(v[i].z, w.z, fmaf(v[i].w, w.w, acc[o]))));
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 3; ++o) acc[o] = warp_sum(acc[o]);
    if (lane < 3) {
        int j = (int)(m % p.J);
        long long sf = m / p.J + (long long)p.s0 * p.F;     // global (s*F + f)
        int g = p.part_joints[j];
        float y = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : acc[2]) + p.bh[lane];
        p.pred[((size_t)sf * p.num_kps + g) * 3 + lane] = y;
    }
    }
}

int launch_head(const HeadParams& p, cudaStream_t st) {
    if (p.M == 0) return 0;
    const int wpb = 8;
    if (p.C % 4 != 0 || p.C > 384) {
        set_last_error("head: C=%d unsupported (multiple of 4, <= 384)", p.C);
        return -1;
    }
    if (p.C <= 256)
        PAFUSE_CUDA_OK(launch_chain(head_kernel<2>, dim3(persistent_blocks<head_kernel<2>>(p.M, wpb)), dim3(wpb * 32), 0, st, 1, p));
    else
        PAFUSE_CUDA_OK(launch_chain(head_kernel<3>, dim3(persistent_blocks<head_kernel<3>>(p.M, wpb)), dim3(wpb * 32), 0, st, 1, p));
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ DDIM step (x0 / eps / update)
// diffusionpose.py:211-225 (un-flip, average, scale, clamp, x0->eps in fp64) and
// :302-312 (img update, fp32, separate mul/add roundings like eager PyTorch).
// One thread per (r, f, joint); the three coordinates are handled together.
__global__ void ddim_step_kernel(DdimParams p) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)p.R * p.F * p.num_kps;
    if (idx >= total) return;
    int g = (int)(idx % p.num_kps);
    long long rf = idx / p.num_kps;                 // r*F + f
    int f = (int)(rf % p.F);
    long long r = rf / p.F;
    long long b = r / p.H;
    int h = (int)(r % p.H);
    const float* po = p.pred + (size_t)idx * 3;
    float x0v[3];
    if (p.flip) {
        const float* pf = p.pred + ((size_t)((long long)p.R * p.F + rf) * p.num_kps + p.flip_perm[g]) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = po[c];
            float bflip = c == 0 ? -pf[c] : pf[c];
            x0v[c] = __fdiv_rn(__fadd_rn(a, bflip), 2.0f);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) x0v[c] = po[c];
    }
    float* x0o = p.x0_out + (size_t)b * p.x0_batch_stride + (((size_t)h * p.F + f) * p.num_kps + g) * 3;
    float* img = p.img + (size_t)idx * 3;
    const float* nz = p.noise ? p.noise + (size_t)idx * 3 : nullptr;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x0 = __fmul_rn(x0v[c], p.scale);
        x0 = clamp_keep_nan(x0, -p.clamp, p.clamp);
        x0o[c] = x0;
        if (p.last) {
            img[c] = x0;
        } else {
            double e64 = __ddiv_rn(__dsub_rn(__dmul_rn(p.sqrt_recip, (double)img[c]), (double)x0), p.sqrt_recipm1);
            if (p.flip) {
                float eps = (float)e64;
                img[c] = __fadd_rn(__fadd_rn(__fmul_rn(x0, p.sqrt_an), __fmul_rn(p.c, eps)), __fmul_rn(p.sigma, nz[c]));
            } else {
                // non-TTA sampler keeps eps in fp64 and casts the sum (diffusionpose.py:189, :265-268)
                double t1 = (double)__fmul_rn(x0, p.sqrt_an);
                double t2 = __dmul_rn(p.c64, e64);
                double t3 = (double)__fmul_rn(p.sigma, nz[c]);
                img[c] = (float)__dadd_rn(__dadd_rn(t1, t2), t3);
            }
        }
    }
}

int launch_ddim_step(const DdimParams& p, cudaStream_t st) {
    long long total = (long long)p.R * p.F * p.num_kps;
    if (total == 0) return 0;
    unsigned blocks = (unsigned)((total + 255) / 256);
    ddim_step_kernel<<<blocks, 256, 0, st>>>(p);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ part re-assembly
// common/utils.py:113-126 with center_pose_at_root(revert=True) (:79-92):
//   out[root] = (-x[root]) + x[root] = +0.0 ;  out[j] = x[j] + x[conn(part of j)]
// The reference also negates the connection rows of its INPUT in place (offset is a
// view); that side effect is a separate launch (negate_rows) ordered after this one.
__global__ void reassemble_kernel(const float* __restrict__ in, float* __restrict__ out,
                                  const int* __restrict__ conn_of_joint, long long poses, int num_kps) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= poses * num_kps) return;
    int g = (int)(idx % num_kps);
    long long pose = idx / num_kps;
    int r = conn_of_joint[g];
    const float* x = in + (size_t)idx * 3;
    float* oo = out + (size_t)idx * 3;
    if (r < 0) {                      // joint belongs to no re-assembled part: stays zero (zeros_like)
        oo[0] = oo[1] = oo[2] = 0.f;
        return;
    }
    const float* xr = in + ((size_t)pose * num_kps + r) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) oo[c] = (g == r) ? __fadd_rn(-x[c], x[c]) : __fadd_rn(x[c], xr[c]);
}

int launch_reassemble(const float* in, float* out, const int* conn_of_joint, long long poses, int num_kps,
                      cudaStream_t st) {
    long long total = poses * num_kps;
    if (total == 0) return 0;
    reassemble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, conn_of_joint, poses, num_kps);
    PAFUSE_LAUNCH_OK();
    return 0;
}

__global__ void negate_rows_kernel(float* __restrict__ x, const int* __restrict__ rows, int nrows, long long poses,
                                   int num_kps) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= poses * nrows * 3) return;
    int c = (int)(idx % 3);
    long long t = idx / 3;
    int r = rows[t % nrows];
    long long pose = t / nrows;
    float* v = x + ((size_t)pose * num_kps + r) * 3 + c;
    *v = -*v;
}

int launch_negate_rows(float* x, const int* rows, int nrows, long long poses, int num_kps, cudaStream_t st) {
    long long total = poses * nrows * 3;
    if (total == 0) return 0;
    negate_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, rows, nrows, poses, num_kps);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ projection + aggregation
// camera.py:30-60 restated with the same operation order (all roundings explicit).
__device__ __forceinline__ void project_point(const float X[3], const float* __restrict__ cam, float uv[2]) {
    float xx0 = clamp_keep_nan(__fdiv_rn(X[0], X[2]), -1.f, 1.f);
    float xx1 = clamp_keep_nan(__fdiv_rn(X[1], X[2]), -1.f, 1.f);
    float r2 = __fadd_rn(__fmul_rn(xx0, xx0), __fmul_rn(xx1, xx1));
    float r4 = __fmul_rn(r2, r2);
    float r6 = __fmul_rn(r4, r2);
    float ksum = __fadd_rn(__fadd_rn(__fmul_rn(cam[4], r2), __fmul_rn(cam[5], r4)), __fmul_rn(cam[6], r6));
    float radial = __fadd_rn(1.f, ksum);
    float tan = __fadd_rn(__fmul_rn(cam[7], xx0), __fmul_rn(cam[8], xx1));
    float rt = __fadd_rn(radial, tan);
    float x0 = __fadd_rn(__fmul_rn(xx0, rt), __fmul_rn(cam[7], r2));
    float x1 = __fadd_rn(__fmul_rn(xx1, rt), __fmul_rn(cam[8], r2));
    uv[0] = __fadd_rn(__fmul_rn(cam[0], x0), cam[2]);
    uv[1] = __fadd_rn(__fmul_rn(cam[1], x1), cam[3]);
}

__global__ void project_kernel(const float* __restrict__ X, const float* __restrict__ cam, float* __restrict__ out,
                               long long npts, long long pts_per_cam) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    float x[3] = {X[i * 3], X[i * 3 + 1], X[i * 3 + 2]};
    float uv[2];
    project_point(x, cam + (i / pts_per_cam) * 9, uv);
    out[i * 2] = uv[0];
    out[i * 2 + 1] = uv[1];
}

int launch_project(const float* X, const float* cam, float* out, long long npts, long long pts_per_cam,
                   cudaStream_t st) {
    if (npts == 0) return 0;
    project_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(X, cam, out, npts, pts_per_cam);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// One thread per (b,k,f,j): loop over the H hypotheses once, keeping the running
// sum (P-Agg, loss.py:68-70) and the first minimum of the 2D reprojection error
// (J-Agg, loss.py:101-108 / visualization.py:453-463; reprojection main_h3wb.py:336-342).
__global__ void aggregate_kernel(AggParams p) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)p.B * p.K * p.F * p.J;
    if (idx >= total) return;
    int j = (int)(idx % p.J);
    long long t = idx / p.J;
    int f = (int)(t % p.F);
    t /= p.F;
    int k = (int)(t % p.K);
    long long b = t / p.K;
    const float* tr = p.traj ? p.traj + ((size_t)b * p.F + f) * 3 : nullptr;
    const float* tgt = p.x2d + (((size_t)b * p.F + f) * p.J + j) * 2;
    const float* cam = p.cam + (p.cam_per_clip ? b * 9 : 0);
    float best = 0.f, bx[3] = {0.f, 0.f, 0.f}, sum[3] = {0.f, 0.f, 0.f};
    int besth = 0;
    for (int h = 0; h < p.H; ++h) {
        const float* x = p.pred + (((((size_t)b * p.K + k) * p.H + h) * p.F + f) * p.J + j) * 3;
        float xv[3] = {x[0], x[1], x[2]};
        float xa[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            xa[c] = tr ? __fadd_rn(xv[c], tr[c]) : xv[c];
            sum[c] = h == 0 ? xv[c] : __fadd_rn(sum[c], xv[c]);
        }
        float uv[2];
        project_point(xa, cam, uv);
        float d0 = __fsub_rn(uv[0], tgt[0]), d1 = __fsub_rn(uv[1], tgt[1]);
        float err = __fsqrt_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)));
        if (p.reproj) {
            float* ro = p.reproj + (((((size_t)b * p.K + k) * p.H + h) * p.F + f) * p.J + j) * 2;
            ro[0] = uv[0];
            ro[1] = uv[1];
        }
        if (h == 0 || err < best) {                 // strict '<' keeps the first minimum
            best = err;
            besth = h;
            bx[0] = xv[0];
            bx[1] = xv[1];
            bx[2] = xv[2];
        }
    }
    float hf = (float)p.H;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        p.jagg[(size_t)idx * 3 + c] = bx[c];
        p.pagg[(size_t)idx * 3 + c] = __fdiv_rn(sum[c], hf);
    }
    if (p.select) p.select[idx] = besth;
}

int launch_aggregate(const AggParams& p, cudaStream_t st) {
    long long total = (long long)p.B * p.K * p.F * p.J;
    if (total == 0) return 0;
    aggregate_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ caller-side preparation
// eval_data_prepare (main_h3wb.py:122-154 / in_the_wild/utils.py:279-320) fused with the flip-TTA input
// construction of the callers (main_h3wb.py:268-270, in_the_wild/utils.py:340-342):
//   clips[i, f]      = seq[src(i, f)]            src = i*F + f for i < n-1; the last clip is the last F frames of the
//                                                sequence, a sequence shorter than F is padded by repeating its last frame
//   clips_flip[i, f, j] = (-x, y) of seq[src(i, f), flip_perm[j]]
__global__ void prepare_clips_kernel(const float* __restrict__ seq, long long T, int F, int J,
                                     const int* __restrict__ flip_perm, float* __restrict__ clips,
                                     float* __restrict__ clips_flip, long long n_clips) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_clips * F * J) return;
    const int j = (int)(idx % J);
    long long t = idx / J;
    const int f = (int)(t % F);
    const long long i = t / F;
    long long src;
    if (i < n_clips - 1) {
        src = i * F + f;
    } else {
        src = (T > F ? T - F : 0) + f;
        if (src > T - 1) src = T - 1;
    }
    const float2 v = *reinterpret_cast<const float2*>(seq + ((size_t)src * J + j) * 2);
    *reinterpret_cast<float2*>(clips + (size_t)idx * 2) = v;
    if (clips_flip) {
        const float2 w = *reinterpret_cast<const float2*>(seq + ((size_t)src * J + flip_perm[j]) * 2);
        *reinterpret_cast<float2*>(clips_flip + (size_t)idx * 2) = make_float2(-w.x, w.y);
    }
}

int launch_prepare_clips(const float* seq, long long T, int F, int J, const int* flip_perm, float* clips,
                         float* clips_flip, long long n_clips, cudaStream_t st) {
    const long long total = n_clips * F * J;
    if (total == 0) return 0;
    prepare_clips_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(seq, T, F, J, flip_perm, clips, clips_flip, n_clips);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// Clips back to the sequence (in_the_wild/h3wb_diffusion.py:119-133): pred [n,K,H,F,J,3] -> out [K,H,T,J,3]; frame t
// comes from clip t / F while t is inside a full clip, the remaining T mod F frames are the LAST ones of the last clip.
__global__ void stitch_clips_kernel(const float* __restrict__ pred, float* __restrict__ out, long long n_clips, int K,
                                    int H, int F, int J, long long T) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)K * H * T * J;
    if (idx >= total) return;
    const int j = (int)(idx % J);
    long long r = idx / J;
    const long long t = r % T;
    r /= T;
    const int h = (int)(r % H);
    const int k = (int)(r / H);
    const long long full = T / F;
    long long clip;
    int f;
    if (t < full * F) {
        clip = t / F;
        f = (int)(t % F);
    } else {
        clip = n_clips - 1;
        f = F - (int)(T - t);
    }
    const float* src = pred + (((((size_t)clip * K + k) * H + h) * F + f) * J + j) * 3;
    float* dst = out + (size_t)idx * 3;
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
}

int launch_stitch_clips(const float* pred, float* out, long long n_clips, int K, int H, int F, int J, long long T,
                        cudaStream_t st) {
    const long long total = (long long)K * H * T * J;
    if (total == 0) return 0;
    stitch_clips_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pred, out, n_clips, K, H, F, J, T);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// OpenPifPaf detections -> normalised H3WB input (in_the_wild/h3wb_diffusion.py:64-77, common/camera.py:7-11):
// raw [T, J-1, 3] (x, y, confidence in pixels) -> kp [T, J, 2]: joints 1.. are the detections, joint 0 is the mean of
// joints 12 and 13 (fp32), then X / w * 2 - [1, h / w] with the subtraction in fp64 like numpy (float32 array minus
// a python-float list) and one rounding to fp32 (the callers' astype('float32')).
__global__ void keypoints_kernel(const float* __restrict__ raw, float* __restrict__ kp, long long T, int J, float w,
                                 double hw) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * J) return;
    const int j = (int)(idx % J);
    const long long t = idx / J;
    const float* fr = raw + (size_t)t * (J - 1) * 3;
    float x, y;
    if (j == 0) {
        x = __fdiv_rn(__fadd_rn(fr[11 * 3], fr[12 * 3]), 2.0f);          // joints 12, 13 = detections 11, 12
        y = __fdiv_rn(__fadd_rn(fr[11 * 3 + 1], fr[12 * 3 + 1]), 2.0f);
    } else {
        x = fr[(j - 1) * 3];
        y = fr[(j - 1) * 3 + 1];
    }
    const float nx = __fmul_rn(__fdiv_rn(x, w), 2.0f), ny = __fmul_rn(__fdiv_rn(y, w), 2.0f);
    kp[(size_t)idx * 2] = (float)__dsub_rn((double)nx, 1.0);
    kp[(size_t)idx * 2 + 1] = (float)__dsub_rn((double)ny, hw);
}

int launch_keypoints(const float* raw, float* kp, long long T, int J, int w, int h, cudaStream_t st) {
    if (T == 0) return 0;
    const long long total = T * J;
    keypoints_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(raw, kp, T, J, (float)w, (double)h / (double)w);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ GT-dependent multi-hypothesis metrics
// The four whole-body MPJPE protocols evaluate() accumulates (main_h3wb.py:344-349) in one pass over the predictions:
//   J-Best  mpjpe_diffusion_all_min(mean_pos=False)  common/loss.py:53-66   per joint min over hypotheses of |p - g|
//   P-Agg   mpjpe_diffusion_all_min(mean_pos=True)   common/loss.py:68-76   |mean_h p - g|
//   J-Agg   mpjpe_diffusion_reproj                   common/loss.py:90-112  |p - g| of the hypothesis with the smallest 2D
//                                                                           reprojection error (first minimum)
//   P-Best  mpjpe_diffusion(mean_pos=False)          common/loss.py:114-146 per hypothesis mean of the root-centred error
//                                                                           (the host takes the min over hypotheses)
// One thread per (b, k, f, j); a block shares (b, k) and adds its partial sums (fp64) to out[k][0..2] and out[k][3+h].
__global__ void metrics_kernel(AggParams p, const float* __restrict__ target, const float* __restrict__ reproj_in,
                               double* __restrict__ out) {
    const int bk = blockIdx.y;
    const int k = bk % p.K;
    const long long b = bk / p.K;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;          // (f, j)
    const int FJ = p.F * p.J;
    const bool live = item < FJ;
    extern __shared__ double red[];                                   // [(3 + H)][warps]
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float jbest = 0.f, jagg = 0.f, pagg = 0.f;
    const int f = live ? item / p.J : 0, j = live ? item % p.J : 0;
    const float* g = target + (((size_t)b * p.F + f) * p.J + j) * 3;
    const float* g0 = target + (((size_t)b * p.F + f) * p.J) * 3;    // root joint of the frame
    const float* tr = p.traj ? p.traj + ((size_t)b * p.F + f) * 3 : nullptr;
    const bool has2d = p.x2d != nullptr;                            // without a 2D target the J-Agg column is not meaningful (0)
    const float* tgt2 = has2d ? p.x2d + (((size_t)b * p.F + f) * p.J + j) * 2 : nullptr;
    const float* cam = p.cam ? p.cam + (p.cam_per_clip ? b * 9 : 0) : nullptr;
    float gv[3] = {0.f, 0.f, 0.f}, gc[3] = {0.f, 0.f, 0.f}, sum[3] = {0.f, 0.f, 0.f};
    if (live) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gv[c] = g[c];
            gc[c] = __fsub_rn(g[c], g0[c]);
        }
    }
    float best2d = 0.f;
    for (int h = 0; h < p.H; ++h) {
        float e_c = 0.f;
        if (live) {
            const size_t base = ((((size_t)b * p.K + k) * p.H + h) * p.F + f) * p.J;
            const float* x = p.pred + (base + j) * 3;
            const float* x0 = p.pred + base * 3;
            float d[3], dc[3], xa[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xv = x[c];
                d[c] = __fsub_rn(xv, gv[c]);
                dc[c] = __fsub_rn(__fsub_rn(xv, x0[c]), gc[c]);
                xa[c] = tr ? __fadd_rn(xv, tr[c]) : xv;
                sum[c] = h == 0 ? xv : __fadd_rn(sum[c], xv);
            }
            const float e = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
            e_c = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dc[0], dc[0]), __fmul_rn(dc[1], dc[1])), __fmul_rn(dc[2], dc[2])));
            if (h == 0 || e < jbest) jbest = e;
            if (has2d) {
                float uv[2];
                if (reproj_in) {
                    const float* r2 = reproj_in + (base + j) * 2;
                    uv[0] = r2[0];
                    uv[1] = r2[1];
                } else {
                    project_point(xa, cam, uv);
                }
                const float d0 = __fsub_rn(uv[0], tgt2[0]), d1 = __fsub_rn(uv[1], tgt2[1]);
                const float e2 = __fsqrt_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)));
                if (h == 0 || e2 < best2d) {
                    best2d = e2;
                    jagg = e;
                }
            }
        }
        // per-hypothesis sum of the root-centred error over the block
        double v = (double)e_c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[(3 + h) * nwarp + warp] = v;
    }
    if (live) {
        const float hf = (float)p.H;
        float m[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) m[c] = __fsub_rn(__fdiv_rn(sum[c], hf), gv[c]);
        pagg = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], m[0]), __fmul_rn(m[1], m[1])), __fmul_rn(m[2], m[2])));
    }
    double v3[3] = {(double)jbest, (double)pagg, (double)jagg};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double v = v3[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[q * nwarp + warp] = v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 3 + p.H; q += blockDim.x) {
        double v = 0.0;
        for (int w = 0; w < nwarp; ++w) v += red[q * nwarp + w];
        atomicAdd(out + (size_t)k * (3 + p.H) + q, v);
    }
}

int launch_metrics(const AggParams& p, const float* target, const float* reproj_in, double* out, cudaStream_t st) {
    if ((long long)p.B * p.K == 0) return 0;
    const int threads = 256;
    dim3 grid((unsigned)((p.F * p.J + threads - 1) / threads), (unsigned)(p.B * p.K));
    const size_t smem = (size_t)(3 + p.H) * (threads / 32) * sizeof(double);
    metrics_kernel<<<grid, threads, smem, st>>>(p, target, reproj_in, out);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// Part-based variants of P-Best and P-Agg (common/loss.py:114-146 and :36-88 with part_based=True; evaluate() calls both,
// main_h3wb.py:350-362): prediction and target are centred PER PART first (center_pose_parts, common/utils.py:95-110:
// joint j of part p minus the part's root joint), then
//   out[k][h][p] (h < H)  = sum over (b, f, j in p) of |pc - gc| of hypothesis h      (P-Best part-based: the host takes the
//                                                                                    hypothesis with the smallest total)
//   out[k][H][p]          = sum over (b, f, j in p) of |mean_h pc - gc|               (P-Agg part-based)
// One thread per (b, k, f, j); block partial sums (fp64) in shared memory, then one atomicAdd per (h, part) and block.
__global__ void metrics_parts_kernel(AggParams p, const float* __restrict__ target, const int* __restrict__ part_of_joint,
                                     const int* __restrict__ root_of_joint, int n_parts, double* __restrict__ out) {
    const int bk = blockIdx.y;
    const int k = bk % p.K;
    const long long b = bk / p.K;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;          // (f, j)
    const int FJ = p.F * p.J;
    extern __shared__ double red[];                                   // [(H + 1)][n_parts]
    const int nred = (p.H + 1) * n_parts;
    for (int i = threadIdx.x; i < nred; i += blockDim.x) red[i] = 0.0;
    __syncthreads();
    const int f = item < FJ ? item / p.J : 0, j = item < FJ ? item % p.J : 0;
    const int part = item < FJ ? part_of_joint[j] : -1;
    if (part >= 0) {
        const int r = root_of_joint[j];
        const float* g = target + (((size_t)b * p.F + f) * p.J) * 3;
        float gc[3], sum[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) gc[c] = __fsub_rn(g[j * 3 + c], g[r * 3 + c]);
        for (int h = 0; h < p.H; ++h) {
            const float* x = p.pred + (((((size_t)b * p.K + k) * p.H + h) * p.F + f) * p.J) * 3;
            float d[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float pc = __fsub_rn(x[j * 3 + c], x[r * 3 + c]);
                d[c] = __fsub_rn(pc, gc[c]);
                sum[c] = h == 0 ? pc : __fadd_rn(sum[c], pc);
            }
            const float e = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
            atomicAdd(&red[h * n_parts + part], (double)e);
        }
        const float hf = (float)p.H;
        float m[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) m[c] = __fsub_rn(__fdiv_rn(sum[c], hf), gc[c]);
        const float e = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], m[0]), __fmul_rn(m[1], m[1])), __fmul_rn(m[2], m[2])));
        atomicAdd(&red[p.H * n_parts + part], (double)e);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nred; i += blockDim.x) atomicAdd(out + (size_t)k * nred + i, red[i]);
}

int launch_metrics_parts(const AggParams& p, const float* target, const int* part_of_joint, const int* root_of_joint,
                         int n_parts, double* out, cudaStream_t st) {
    if ((long long)p.B * p.K == 0) return 0;
    const int threads = 256;
    dim3 grid((unsigned)((p.F * p.J + threads - 1) / threads), (unsigned)(p.B * p.K));
    const size_t smem = (size_t)(p.H + 1) * n_parts * sizeof(double);
    metrics_parts_kernel<<<grid, threads, smem, st>>>(p, target, part_of_joint, root_of_joint, n_parts, out);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------ counter-based Gaussian noise (multi-GPU sampler draws)
// The sampler's draws (diffusionpose.py:283, :308) as a pure function of (seed, draw, element index): element e of draw
// `stream_id` is Box-Muller of the Philox4x32-10 block with counter (e >> 1, stream_id) and key = seed, cosine branch,
// using words (0,1) for even e and (2,3) for odd e.  Any rank can therefore generate exactly its slice of the global
// (B,H,F,J,3) tensor -- no rank draws the whole tensor to keep a slice, and the union of the shards is bit-identical to
// the single-GPU draw whatever the sharding.  The transform runs in fp64 (one rounding to fp32 at the end), which makes
// the values reproducible to the last bit against a numpy restatement (tests/test_host_logic.py).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void randn_philox_kernel(float* __restrict__ out, unsigned long long seed, unsigned long long stream_id,
                                    long long base, long long rows, long long row_len, long long row_stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * row_len) return;
    const long long r = i / row_len;
    const unsigned long long e = (unsigned long long)(base + r * row_stride + (i - r * row_len));
    const unsigned long long ctr = e >> 1;
    uint32_t w[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), w);
    const uint32_t a = w[2 * (e & 1)], b2 = w[2 * (e & 1) + 1];
    const double u1 = ((double)a + 0.5) * (1.0 / 4294967296.0);      // (0, 1)
    const double u2 = ((double)b2 + 0.5) * (1.0 / 4294967296.0);
    out[i] = (float)(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
}

int launch_randn_philox(float* out, unsigned long long seed, unsigned long long stream_id, long long base, long long rows,
                        long long row_len, long long row_stride, cudaStream_t st) {
    const long long total = rows * row_len;
    if (total == 0) return 0;
    randn_philox_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, seed, stream_id, base, rows, row_len, row_stride);
    PAFUSE_LAUNCH_OK();
    return 0;
}

// max |w| of a GEMM weight tensor, as the bit pattern of a non-negative float (monotone under atomicMax on unsigned);
// a NaN compares as larger than everything finite, which is what the range check wants.
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, unsigned int* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned int m = 0u;
    for (; i < n; i += stride) m = max(m, __float_as_uint(w[i]) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

int launch_absmax(const float* w, size_t n, unsigned int* out, cudaStream_t st) {
    if (n == 0) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    absmax_kernel<<<blocks, 256, 0, st>>>(w, n, out);
    PAFUSE_LAUNCH_OK();
    return 0;
}

}  // namespace pafuse
