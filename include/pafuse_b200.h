/*
 * pafuse_b200 -- C ABI of the B200-native PAFUSE denoising inference path.
 *
 * The reference (valeoai/PAFUSE) is pure PyTorch and has no FFI; the boundary this
 * library replaces is the eval branch of the nn.Module contract of `D3DP` /
 * `MixSTE2` plus three free functions.  Each entry point cites the reference
 * interface it stands in for (paths relative to the reference root).  The Python
 * host module `pafuse_b200/diffusionpose.py` binds these symbols with ctypes and
 * keeps the reference's constructor / forward signatures.
 *
 * Conventions
 *   - plain C types only; all tensor arguments are DEVICE pointers to contiguous
 *     fp32 data unless stated otherwise; the caller owns them.
 *   - every call enqueues work on the caller's CUDA stream (`stream` is a
 *     cudaStream_t passed as void*) and returns without synchronising.
 *   - return value: 0 = ok, negative = error (PAFUSE_E_*); `pafuse_last_error()`
 *     returns a thread-local description.  Nothing throws across the ABI.
 *   - a context is bound to the device current at `pafuse_create` and is not
 *     thread-safe.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with PAFUSE_E_CUDA.
 */
#ifndef PAFUSE_B200_H
#define PAFUSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAFUSE_MAX_PARTS 4
#define PAFUSE_E_ARG (-1)
#define PAFUSE_E_CUDA (-2)
#define PAFUSE_E_STATE (-3)

typedef struct pafuse_ctx pafuse_ctx;

/* Static description of the part-based denoiser.  Mirrors what D3DP.__init__ derives
 * from `args` and `dataset` (common/diffusionpose.py:59-153): frames, num_kps,
 * depth (args.model.dep), heads (8), the part table {body,face,hands} with its
 * channel widths (:141) and joint-index lists (:72-83), the flip-TTA joint
 * permutation built from joints_left/right (:197-198) and args.ft2d.scale. */
typedef struct pafuse_config {
    int32_t frames;                              /* F, args.model.number_of_frames (27) */
    int32_t num_kps;                             /* args.data.num_kps (134) */
    int32_t depth;                               /* args.model.dep (8) */
    int32_t heads;                               /* 8 */
    int32_t num_parts;                           /* 3 */
    int32_t part_channels[PAFUSE_MAX_PARTS];     /* 384 / 224 / 256 */
    int32_t part_num_joints[PAFUSE_MAX_PARTS];   /* 24 / 68 / 42 */
    const int32_t* part_joints[PAFUSE_MAX_PARTS];/* host arrays: whole-body joint ids, concat order */
    const int32_t* flip_perm;                    /* host array [num_kps]: source joint under L/R swap */
    float scale;                                 /* args.ft2d.scale */
    int32_t max_seqs;                            /* sequences (clip x hypothesis x {orig,flip}) per workspace pass */
} pafuse_config;

const char* pafuse_last_error(void);
const char* pafuse_version(void);

/* Number of CUDA kernels this library has launched on the calling thread (bench.py's gpu_launches). */
int64_t pafuse_launch_count(void);

int pafuse_create(const pafuse_config* cfg, pafuse_ctx** out);
void pafuse_destroy(pafuse_ctx* ctx);

/* Weights: replaces nn.Module.load_state_dict for `pose_estimator.<part>.<name>`
 * (keys/shapes of common/mixste.py:141-210).  `name` is the sub-key after the part,
 * e.g. "STEblocks.3.attn.qkv.weight"; `data` is fp32, host or device (`on_device`).
 * Unknown names return PAFUSE_E_ARG.  Call pafuse_commit_weights once all tensors
 * are set: it derives the fp16 hi/lo operand copies the tensor-core GEMMs read. */
int pafuse_set_weight(pafuse_ctx* ctx, int32_t part, const char* name, const float* data, int64_t numel,
                      int32_t on_device);
int pafuse_commit_weights(pafuse_ctx* ctx, void* stream);

/* pred_parts: D3DP.pred_parts + split_data (common/diffusionpose.py:163-172,328-335)
 * = MixSTE2.forward per part (common/mixste.py:278-298), results concatenated on the
 * joint axis.  x2d [B,F,num_kps,2], x3d [B,H,F,num_kps,3] (already clamped/scaled
 * x_t), sinus = concatenated sinusoidal timestep embeddings of the parts
 * (mixste.py:132-139; sum(part_channels) floats, device) -> out [B,H,F,num_kps,3]. */
int pafuse_pred_parts(pafuse_ctx* ctx, const float* x2d, const float* x3d, const float* sinus, float* out, int32_t B,
                      int32_t H, void* stream);

/* One DDIM step: D3DP.model_predictions_fliping (flip != 0, diffusionpose.py:192-225)
 * or model_predictions (flip == 0, :174-190) followed by the state update of
 * ddim_sample[_flip] (:298-312).
 *   img      [B,H,F,num_kps,3]  sampler state, updated in place
 *   noise    [B,H,F,num_kps,3]  the randn_like draw of this step (ignored when last != 0)
 *   x0_out   base pointer of preds_all[:, k]; clip b is written at x0_out + b*x0_batch_stride
 *   sqrt_recip, sqrt_recipm1    fp64 buffers at t (:119-120); sqrt_an, c, sigma: fp32 casts of the
 *   fp64 scalars of :302-306 (c64 = the fp64 value of c, used by the non-flip sampler)        */
int pafuse_ddim_step(pafuse_ctx* ctx, const float* x2d, const float* x2d_flip, const float* sinus, float* img,
                     const float* noise, float* x0_out, int64_t x0_batch_stride, int32_t B, int32_t H, int32_t flip,
                     int32_t last, double sqrt_recip, double sqrt_recipm1, double c64, float sqrt_an, float c,
                     float sigma, void* stream);

/* wb_pose_from_parts (common/utils.py:113-126): pose [poses,num_kps,3] -> out.
 * conn_of_joint: host array [num_kps], connection joint of the part each joint belongs to
 * (body -> 0, face -> 1, left hand -> 10, right hand -> 11; -1 = untouched/zero).
 * mutate_input != 0 reproduces the reference's in-place negation of the connection rows of `pose`. */
int pafuse_wb_pose_from_parts(pafuse_ctx* ctx, float* pose, float* out, const int32_t* conn_of_joint, int64_t poses,
                              int32_t mutate_input, void* stream);

/* project_to_2d (common/camera.py:30-60): X [n_cams, pts_per_cam, 3], cam [n_cams,9] -> out [.,.,2]. */
int pafuse_project_to_2d(pafuse_ctx* ctx, const float* X, const float* cam, float* out, int64_t n_cams,
                         int64_t pts_per_cam, void* stream);

/* Multi-hypothesis aggregation: reprojection (main_h3wb.py:336-342), J-Agg select
 * (common/loss.py:101-108, pose form common/visualization.py:453-463) and P-Agg mean
 * (loss.py:68-70).  pred [B,K,H,F,J,3] whole-body, traj [B,F,1,3] or NULL, cam [1,9]
 * (cam_per_clip == 0) or [B,9], x2d [B,F,J,2] -> jagg,pagg [B,K,F,J,3]; select [B,K,F,J] int32
 * and reproj [B,K,H,F,J,2] are optional (NULL to skip). */
int pafuse_aggregate(pafuse_ctx* ctx, const float* pred, const float* traj, const float* cam, int32_t cam_per_clip,
                     const float* x2d, float* jagg, float* pagg, int32_t* select, float* reproj, int32_t B, int32_t K,
                     int32_t H, void* stream);

/* GT-dependent multi-hypothesis MPJPE protocols of evaluate() (main_h3wb.py:344-349; common/loss.py:36-146), one pass:
 * pred [B,K,H,F,J,3] whole-body, target [B,F,J,3]; the 2D error of J-Agg uses `reproj` [B,K,H,F,J,2] when given, else
 * the projection of pred + traj with cam as in pafuse_aggregate.  sums (device, fp64) [K][3+H]: per sampling step the
 * SUMS over (b,f,j) of the J-Best, P-Agg and J-Agg errors and, per hypothesis, of the root-centred error (P-Best = min
 * over hypotheses after the division); divide by B*F*J. */
int pafuse_mpjpe_metrics(pafuse_ctx* ctx, const float* pred, const float* target, const float* traj, const float* cam,
                         int32_t cam_per_clip, const float* x2d, const float* reproj, double* sums, int32_t B, int32_t K,
                         int32_t H, void* stream);

/* Part-based variants of the same protocols (common/loss.py:114-146 `mpjpe_diffusion(part_based=True)` and :36-88
 * `mpjpe_diffusion_all_min(mean_pos=True, part_based=True)`, both called by evaluate(), main_h3wb.py:350-362):
 * prediction and target are centred per part first (center_pose_parts, common/utils.py:95-110).
 * part_of_joint / root_of_joint: host arrays [num_kps] -- index of the part a joint belongs to (-1: none) and the root
 * joint of that part (dataset.parts_joint_indices / dataset.root_indices).  sums (device, fp64) [K][H+1][n_parts]:
 * rows h < H hold, per part, the SUM over (b,f,j in part) of the part-centred error of hypothesis h; row H the same for
 * the mean pose over the hypotheses.  The host divides by the joint counts and picks the best hypothesis. */
int pafuse_mpjpe_metrics_parts(pafuse_ctx* ctx, const float* pred, const float* target, const int32_t* part_of_joint,
                               const int32_t* root_of_joint, int32_t n_parts, double* sums, int32_t B, int32_t K, int32_t H,
                               void* stream);

/* The sampler's Gaussian draws (torch.randn / randn_like at common/diffusionpose.py:283,308) as a counter-based
 * generator, so that a rank of a multi-GPU run produces exactly ITS slice of the global (B,H,F,num_kps,3) tensor:
 * out[r*row_len + i] = N(0,1) value of global element `base + r*row_stride + i` of draw number `draw` under `seed`
 * (Philox4x32-10, Box-Muller in fp64).  Clip shards are one row; hypothesis shards are B rows of (h1-h0)*F*num_kps*3
 * elements with stride H*F*num_kps*3.  The union of any sharding is bit-identical to the one-row global draw. */
int pafuse_randn(pafuse_ctx* ctx, uint64_t seed, uint64_t draw, int64_t base, float* out, int64_t rows, int64_t row_len,
                 int64_t row_stride, void* stream);

/* ---- caller-side preparation (the code around the model call in main_h3wb.py / in_the_wild) ----
 *
 * pafuse_prepare_clips: eval_data_prepare (main_h3wb.py:122-154, in_the_wild/utils.py:279-320) fused with the flip-TTA
 * input construction (main_h3wb.py:268-270, in_the_wild/utils.py:340-342).  seq [T,num_kps,2] -> clips
 * [ceil(T/F),F,num_kps,2] (last clip = last F frames; T < F: last frame repeated) and, when clips_flip != NULL, the
 * same clips of the flipped sequence (x negated, left/right joints swapped with the context's flip_perm). */
int pafuse_prepare_clips(pafuse_ctx* ctx, const float* seq, int64_t T, float* clips, float* clips_flip, void* stream);

/* Clips back to one sequence (in_the_wild/h3wb_diffusion.py:119-133): pred [n_clips,K,H,F,num_kps,3] ->
 * out [K,H,T,num_kps,3]; the last T mod F frames are the last frames of the last clip. */
int pafuse_stitch_clips(pafuse_ctx* ctx, const float* pred, int64_t n_clips, int32_t K, int32_t H, int64_t T, float* out,
                        void* stream);

/* OpenPifPaf detections to the model's 2D input (in_the_wild/h3wb_diffusion.py:64-77 + normalize_screen_coordinates,
 * common/camera.py:7-11): raw [T,num_kps-1,3] (x, y, confidence; pixels) -> kp [T,num_kps,2]; joint 0 is the mean of
 * joints 12 and 13. */
int pafuse_keypoints_from_detections(pafuse_ctx* ctx, const float* raw, int64_t T, int32_t width, int32_t height, float* kp,
                                     void* stream);

/* Per-launch device timing for the roofline leg of bench.py: while enabled, every kernel this
 * context launches is bracketed by CUDA events on the launching stream.  pafuse_profile_read
 * synchronises and returns, per category (0 tensor-core GEMM, 1 attention, 2 LayerNorm chain,
 * 3 embed/time-MLP/head, 4 DDIM update, 5 re-assembly + aggregation), the summed device time in
 * ms, the algorithmic work (FLOPs for 0-1, bytes for 2-5) and the launch count.  Enabling again
 * (or disabling) clears the records. */
#define PAFUSE_PROFILE_CATEGORIES 6
int pafuse_profile_enable(pafuse_ctx* ctx, int32_t enable);
int pafuse_profile_read(pafuse_ctx* ctx, double* ms, double* work, int64_t* launches, int32_t ncat);
/* the same records seen as DRAM traffic: per category the algorithmic bytes (operands read once, results written
 * once; for the GEMMs fp16 hi/lo operands, the weight matrix once, and the epilogue's reads and writes) */
int pafuse_profile_read_bytes(pafuse_ctx* ctx, double* bytes, int32_t ncat);

/* ---- unit-level entry points (tests and profiling; same kernels the path uses) ---- */

/* y = x W^T + b through the tcgen05 f16x3 GEMM (use_simt != 0: CUDA-core debug reference).
 * x [M,K] fp32, w [N,K] fp32, b [N]; epilogue 0: y fp32 [M,N]; 1: y = gelu(.) returned as fp32
 * (hi+lo recombined); 2: y += x W^T + b in place. */
int pafuse_linear(pafuse_ctx* ctx, const float* x, const float* w, const float* b, float* y, int64_t M, int32_t N,
                  int32_t K, int32_t epilogue, int32_t use_simt, void* stream);

/* softmax(q k^T / sqrt(hd)) v on qkv [S*F*J, 3C] -> out fp32 [S*F*J, C]; temporal selects the axis. */
int pafuse_attention(pafuse_ctx* ctx, const float* qkv, float* out, int32_t S, int32_t J, int32_t C, int32_t temporal,
                     void* stream);

/* Attention.forward up to (excluding) proj (mixste.py:63-79) the way the path runs it: the qkv GEMM
 * writes fp16 hi/lo head planes from its epilogue, the tcgen05 attention kernel consumes them.
 * x [S*F*J, C] fp32 (the LayerNorm output), w [3C,C], b [3C] -> out fp32 [S*F*J, C]. */
int pafuse_qkv_attention(pafuse_ctx* ctx, const float* x, const float* w, const float* b, float* out, int32_t S,
                         int32_t J, int32_t C, int32_t temporal, void* stream);

/* Mlp.forward of a block with what follows it (mixste.py:37-43, :115, and the norms of :243-273), the way the path
 * runs it for the parts with C <= 256:  x <- x + fc2(GELU(fc1(a)));  g0 != NULL: x <- LN(x; g0, bb0);
 * a_out = LN(x; g1, bb1) (as the fp32 sum of its fp16 hi/lo pair).  fused != 0: one launch (mlp_fused_kernel),
 * else the fc1 and fc2 GEMM launches.  a [M,C], w1 [2C,C], b1 [2C], w2 [C,2C], b2 [C], x [M,C] in/out. */
int pafuse_mlp_block(pafuse_ctx* ctx, const float* a, const float* w1, const float* b1, const float* w2, const float* b2,
                     float* x, const float* g0, const float* bb0, const float* g1, const float* bb1, float* a_out, int64_t M,
                     int32_t C, int32_t fused, void* stream);

/* debugging switch: route the path's GEMMs through the CUDA-core reference kernel */
int pafuse_set_debug_simt_gemm(pafuse_ctx* ctx, int32_t enable);
/* debugging switch: CUDA-core attention kernel instead of the tcgen05 one */
int pafuse_set_debug_simt_attention(pafuse_ctx* ctx, int32_t enable);
/* 1 (default): the LayerNorms that follow the proj / fc2 GEMMs are computed in their epilogues for parts whose
 * channel width fits one output tile (C <= 256); 0: separate LayerNorm launches everywhere */
int pafuse_set_fuse_layernorm(pafuse_ctx* ctx, int32_t enable);
/* 1: Mlp.forward (mixste.py:37-43) of the parts with C <= 256 runs as ONE kernel (fc1 + GELU + fc2 + residual +
 * the following LayerNorms), the hidden activations staying in tensor memory; 0: fc1 and fc2 GEMM launches */
int pafuse_set_fuse_mlp(pafuse_ctx* ctx, int32_t enable);
/* 1: the part denoisers (body / face / hands, independent until the DDIM update: diffusionpose.py:163-172) run
 * side by side, each on its own stream and on a share of the SMs, so that the DRAM-bound kernels of one part
 * overlap the tensor-bound kernels of another; 0 (default; the overlap measured neutral on power-capped B200s):
 * one after the other on the caller's stream.  Results are bit-identical either way.
 * shares: SMs per part (NULL = proportional to J*C). */
int pafuse_set_part_streams(pafuse_ctx* ctx, int32_t enable, const int32_t* shares);
/* Small batches (CPU-plumbing config, the tail batch of a long video, a hypothesis shard of 1-2 clips) are bound by
 * the ~1400 launches of a denoiser pass, not by the device: passes of at most `max_seqs` sequences (default 96;
 * 0 disables) are captured into a CUDA graph per (input pointers, shape) key on their second occurrence and
 * replayed afterwards.  pafuse_graph_replays counts the passes that ran as one graph launch. */
int pafuse_set_graph_max_seqs(pafuse_ctx* ctx, int32_t max_seqs);
int64_t pafuse_graph_replays(pafuse_ctx* ctx);
/* process-wide: 2 (default) = tcgen05 CTA pairs (cta_group::2, 256-row tiles), 1 = lone CTAs */
int pafuse_set_gemm_cta_group(int32_t cta_group);
/* process-wide: 1 (default) = weight-stationary GEMM tiles where the W slice fits in shared memory, 0 = always stream W */
int pafuse_set_gemm_weight_stationary(int32_t enable);

#ifdef __cplusplus
}
#endif
#endif /* PAFUSE_B200_H */
