// Kernel parameter blocks and host-side launch entry points shared by the .cu files.
#pragma once

#include "common.cuh"

namespace pafuse {

struct EmbedParams {
    long long M;              // token rows in this chunk = Sc*F*J
    int C, J, F, H;
    int R;                    // number of (clip,hypothesis) rows; sequences >= R are flip twins
    int s0;                   // first global sequence id of this chunk
    int num_kps;
    int apply_clamp;          // x3d is the raw sampler state: clamp(+-clamp) and divide by scale
    float clamp, scale;
    const float* x2d;         // [B,F,num_kps,2]
    const float* x2d_flip;    // [B,F,num_kps,2] or null
    const float* x3d;         // [R,F,num_kps,3]
    const int* part_joints;   // [J] whole-body joint ids (device)
    const int* flip_perm;     // [num_kps] (device)
    const float *we, *be;     // [C,5], [C]
    const float* spos;        // [J,C]
    const float* temb;        // [C]
    float* x;                 // [M,C] out
    // optional: norm1 of STE block 0 applied to the row in the same pass -> fp16 hi/lo (nullptr: embedding only)
    const float *g1 = nullptr, *b1 = nullptr;
    float eps1 = 1e-6f;
    op_t *out_hi = nullptr, *out_lo = nullptr;
};

struct LnParams {
    long long M;
    int C, J, F;
    float* x;                 // [M,C] in (and out when g0 != null)
    const float *g0, *b0;     // optional shared norm written back to x
    float eps0;
    const float* add_f;       // optional [F,C] added after the first norm (Temporal_pos_embed)
    const float *g1, *b1;     // optional second norm -> fp16 hi/lo
    float eps1;
    op_t *out_hi, *out_lo;
};

struct HeadParams {
    long long M;
    int C, J, F;
    int s0;
    int num_kps;
    const float* x;
    const float *g0, *b0;     // shared Temporal_norm (eps0)
    float eps0;
    const float *g1, *b1;     // head LayerNorm (eps1 = 1e-5)
    float eps1;
    const float *wh, *bh;     // [3,C], [3]
    const int* part_joints;
    float* pred;              // [S,F,num_kps,3]
};

struct DdimParams {
    int R, F, H, num_kps;
    int flip, last;
    const float* pred;        // [2R or R, F, num_kps, 3]
    const int* flip_perm;
    float* img;               // [R,F,num_kps,3] in/out
    const float* noise;       // [R,F,num_kps,3] or null when last
    float* x0_out;            // base of preds_all[:, k]
    long long x0_batch_stride;// elements between consecutive clips in x0_out
    float scale, clamp;
    double sqrt_recip, sqrt_recipm1, c64;
    float sqrt_an, c, sigma;
};

struct AggParams {
    int B, K, H, F, J;
    int cam_per_clip;
    const float* pred;        // [B,K,H,F,J,3]
    const float* traj;        // [B,F,1,3] or null
    const float* cam;         // [1,9] or [B,9]
    const float* x2d;         // [B,F,J,2]
    float* jagg;              // [B,K,F,J,3]
    float* pagg;              // [B,K,F,J,3]
    int* select;              // [B,K,F,J] or null
    float* reproj;            // [B,K,H,F,J,2] or null
};

struct AttnParams {
    const float* qkv;         // [M,3C]
    op_t *out_hi, *out_lo;  // [M,C]
    int S, F, J, C;
    int temporal;             // 0: attend over J inside (s,f); 1: over F inside (s,j)
    float scale;              // head_dim^-0.5, set by launch_attention
};

// q/k/v of one part as per-head planes [which*8 + head][rows_cap][hds], fp16 hi/lo: written by the
// qkv GEMM epilogue (EPI_PLANES), read by attention_tc.cu
struct AttnPlanes {
    op_t* hi;
    op_t* lo;
    long long rows_cap;       // rows per plane (multiple of J)
    int hds;                  // stored head width = head_dim rounded up to 16 (48 / 32 / 32); pad columns are zero
};
inline int attn_head_store(int hd) { return (hd + 15) / 16 * 16; }

enum GemmEpilogue { EPI_F32 = 0, EPI_GELU_SPLIT = 1, EPI_RESID = 2, EPI_PLANES = 3, EPI_RESID_LN = 4 };

// EPI_RESID_LN: x <- x + A W^T + bias, fused with the LayerNorm(s) that follow (see gemm_tcgen05.cu).
//   g0 == nullptr:  x <- v;                               out_hi/lo = LN(v; g1, b1, eps1)
//   g0 != nullptr:  x <- y = LN(v; g0, b0, eps0) [+ add_f[f]];   out_hi/lo = LN(y; g1, b1, eps1)
// Requires the whole row in one tile (N <= 256, N % 32 == 0).
struct GemmLnFuse {
    const float* x = nullptr;            // [M,N] residual stream (read here; written through GemmArgs::out_f32 == x)
    const float *g0 = nullptr, *b0 = nullptr;
    const float* add_f = nullptr;        // [F,N], row f = (m / J) % F
    const float *g1 = nullptr, *b1 = nullptr;
    float eps0 = 1e-6f, eps1 = 1e-6f;
    int J = 1, F = 1;
};

struct GemmArgs {
    const op_t *a_hi, *a_lo;   // [M,K]
    const op_t *w_hi, *w_lo;   // [N,K]
    const float* bias;                  // [N]
    float* out_f32;                     // [M,N]  (EPI_F32: written; EPI_RESID: out += acc + bias)
    op_t *out_hi, *out_lo;     // [M,N]  (EPI_GELU_SPLIT)
    long long M;
    int N, K;
    int epilogue;
    float out_scale = WEIGHT_UNSCALE;   // accumulator scale applied before the bias (weights are stored pre-scaled)
    // EPI_PLANES: N = 24*hds output columns in plane order (column n -> plane n / hds, d = n % hds)
    AttnPlanes planes = {nullptr, nullptr, 0, 0};
    GemmLnFuse ln;                      // EPI_RESID_LN
    int sm_limit = 0;                   // > 0: use at most this many SMs (the parts run side by side on SM shares)
};

// Fused MLP of a block (see mlp_fused_kernel): x <- x + fc2(GELU(fc1(a))) and the LayerNorms of EPI_RESID_LN
struct MlpArgs {
    const op_t *a_hi, *a_lo;   // [M,C]   norm2 output
    const op_t *w1_hi, *w1_lo; // [2C,C]  fc1 weight
    const float* b1;           // [2C]
    const op_t *w2_hi, *w2_lo; // [C,2C]  fc2 weight
    const float* b2;           // [C]
    float* x;                  // [M,C]   residual stream, updated in place (== ln.x)
    op_t *out_hi, *out_lo;     // [M,C]   LayerNorm output for the next GEMM
    long long M;
    int C;
    float out_scale = WEIGHT_UNSCALE;
    GemmLnFuse ln;
    int sm_limit = 0;
};
bool mlp_can_fuse(int C);
int launch_mlp_fused(const MlpArgs& g, cudaStream_t st);

int launch_split_weights(const float* w, op_t* hi, op_t* lo, size_t n, cudaStream_t st);
int launch_time_mlp(const float* sinus, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* temb, int C, cudaStream_t st);
int launch_embed(const EmbedParams& p, cudaStream_t st);
int launch_ln_chain(const LnParams& p, cudaStream_t st);
int launch_head(const HeadParams& p, cudaStream_t st);
int launch_ddim_step(const DdimParams& p, cudaStream_t st);
int launch_reassemble(const float* in, float* out, const int* conn_of_joint, long long poses, int num_kps,
                      cudaStream_t st);
int launch_negate_rows(float* x, const int* rows, int nrows, long long poses, int num_kps, cudaStream_t st);
int launch_project(const float* X, const float* cam, float* out, long long npts, long long pts_per_cam,
                   cudaStream_t st);
int launch_aggregate(const AggParams& p, cudaStream_t st);
int launch_prepare_clips(const float* seq, long long T, int F, int J, const int* flip_perm, float* clips,
                         float* clips_flip, long long n_clips, cudaStream_t st);
int launch_stitch_clips(const float* pred, float* out, long long n_clips, int K, int H, int F, int J, long long T,
                        cudaStream_t st);
int launch_metrics(const AggParams& p, const float* target, const float* reproj_in, double* out, cudaStream_t st);
int launch_keypoints(const float* raw, float* kp, long long T, int J, int w, int h, cudaStream_t st);
int launch_metrics_parts(const AggParams& p, const float* target, const int* part_of_joint, const int* root_of_joint,
                         int n_parts, double* out, cudaStream_t st);
int launch_randn_philox(float* out, unsigned long long seed, unsigned long long stream_id, long long base, long long rows,
                        long long row_len, long long row_stride, cudaStream_t st);
int launch_absmax(const float* w, size_t n, unsigned int* out, cudaStream_t st);
int launch_attention(const AttnParams& p, cudaStream_t st);          // CUDA-core version (debug reference)
int launch_attention_tc(const AttnPlanes& pl, op_t* o_hi, op_t* o_lo, int S, int F, int J, int C, int temporal,
                        cudaStream_t st, int sm_limit = 0);
int launch_qkv_to_planes(const float* qkv, const AttnPlanes& pl, long long M, int C, cudaStream_t st);
// qkv weight [3C,C] / bias [3C] -> plane order with every head padded to hds rows (zero rows / zero bias)
int launch_pack_qkv(const float* w, const float* b, float* wp, float* bp, int C, int hd, int hds, cudaStream_t st);

// tcgen05 GEMM (f16x3 split precision).  Tensor maps are built per call from the raw pointers.
int gemm_init();                                            // resolves cuTensorMapEncodeTiled
int launch_gemm_tcgen05(const GemmArgs& g, cudaStream_t st);
int launch_gemm_simt(const GemmArgs& g, cudaStream_t st);   // debug reference (CUDA cores)
int gemm_pick_block_n(int N);
bool gemm_can_fuse_ln(int N);                               // EPI_RESID_LN needs the row in one tile
void gemm_set_cta_group(int cg);                            // 1: lone CTAs, 2 (default): tcgen05 CTA pairs
void gemm_set_weight_stationary(int on);                    // 1 (default): keep the W slice of an n tile in shared memory when it fits

}  // namespace pafuse
