// C ABI (include/pafuse_b200.h): context, weight store, workspace and the
// orchestration of one denoiser pass / one DDIM step out of the sm_100a kernels.
#include "../../include/pafuse_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace pafuse {

static thread_local char g_err[1024] = "";
thread_local long long g_launch_count = 0;

// Programmatic dependent launch is opt-in (PAFUSE_PDL=1): measured neutral on the power-capped B200s of this pool
// (profiles/r1n_*), and it is suspended while the parts run side by side on SM shares, where early-resident CTAs
// of a dependent kernel would squat on SMs of another part's share.
thread_local int g_pdl_suspend = 0;
int g_switch_epoch = 0;   // bumped by the process-wide GEMM switches: captured passes of every context become stale
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("PAFUSE_PDL");
        return e && atoi(e) != 0;
    }();
    return on && g_pdl_suspend == 0;
}

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct Slot {
    size_t off = 0;      // element offset inside the part arena
    size_t numel = 0;
    bool gemm = false;   // needs a fp16 hi/lo copy
    bool set = false;
    bool derived = false;// filled by pafuse_commit_weights, not by the caller
};

struct Part {
    int C = 0, J = 0;
    std::map<std::string, Slot> slots;
    size_t total = 0;
    float* f32 = nullptr;
    op_t* hi = nullptr;
    op_t* lo = nullptr;
    int* joints_dev = nullptr;
    float* temb = nullptr;

    const float* w(const std::string& n) const { return f32 + slots.at(n).off; }
    const op_t* wh(const std::string& n) const { return hi + slots.at(n).off; }
    const op_t* wl(const std::string& n) const { return lo + slots.at(n).off; }
};

// key of a captured denoiser pass: every pointer and shape the captured launches depend on
struct GraphKey {
    const void *x2d, *x2d_flip, *x3d, *sinus, *pred;
    int apply_clamp, R, H, S_total;
    bool operator<(const GraphKey& o) const {
        return memcmp(this, &o, sizeof(GraphKey)) < 0;
    }
};
struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int seen = 0;                    // direct runs so far (the first one allocates and configures: never captured)
    long long launches = 0;          // kernel nodes of the graph
};

struct Workspace {
    long long rows_x_c = 0;   // capacity in (rows * C) units
    size_t plane_halves = 0;
    float* x = nullptr;
    op_t *a_hi = nullptr, *a_lo = nullptr;
    float* qkv = nullptr;                      // fp32 qkv: only for the CUDA-core debug attention
    op_t *pl_hi = nullptr, *pl_lo = nullptr;   // q/k/v head planes, 24 * rows * hds halves each
    op_t *o_hi = nullptr, *o_lo = nullptr;
    op_t *h_hi = nullptr, *h_lo = nullptr;
};

// Optional per-launch device timing (bench.py's roofline leg): CUDA events on the launching
// stream around every kernel, summed per category when read.
enum ProfCat { CAT_GEMM = 0, CAT_ATTN, CAT_LN, CAT_EMBED_HEAD, CAT_DDIM, CAT_POST, CAT_COUNT };

struct ProfRec {
    int cat;
    double work;          // algorithmic FLOPs (GEMM, attention) or bytes (the rest) of this launch
    double bytes;         // algorithmic DRAM bytes of this launch (operands read once, results written once)
    cudaEvent_t a, b;
};

struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
};

}  // namespace pafuse

using namespace pafuse;

struct pafuse_ctx {
    pafuse_config cfg;
    int device = 0;
    int num_parts = 0;
    Part parts[PAFUSE_MAX_PARTS];
    int* flip_perm_dev = nullptr;
    int* conn_dev = nullptr;         // wb_pose_from_parts tables, resident: uploaded when the caller's table changes
    int* conn_rows_dev = nullptr;
    std::vector<int> conn_host, conn_rows_host;
    int* part_of_joint_dev = nullptr;   // part-based metrics tables, resident the same way
    int* root_of_joint_dev = nullptr;
    std::vector<int> part_of_joint_host, root_of_joint_host;
    unsigned int* absmax_dev = nullptr; // range check of the GEMM weights at commit
    // Small-R path: below graph_max_seqs sequences per pass the ~1400 launches of a denoiser pass cost more on the
    // host than on the device, so a pass is captured once per (inputs, shape) key into a CUDA graph and replayed.
    std::map<GraphKey, GraphEntry> graphs;
    int graph_max_seqs = 96;         // PAFUSE_GRAPH_MAX_SEQS (0 disables); the bench shapes (640 sequences) never capture
    cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the caller's may be the legacy default stream)
    long long graph_replays = 0;
    int graph_epoch = 0;
    Workspace ws[PAFUSE_MAX_PARTS];  // [0] serves every part in the sequential mode; one per part when they run side by side
    // side-by-side mode: the part denoisers are independent until the DDIM update, so each runs on its own stream
    // on a share of the SMs (persistent kernels sized to the share); the HBM-bound kernels of one part (proj, fc2,
    // attention) then overlap the tensor-bound ones of another (qkv, fc1) instead of taking turns on the whole GPU
    bool part_streams = false;       // opt-in (PAFUSE_PART_STREAMS=1 / pafuse_set_part_streams): measured neutral on the
                                     // power-capped B200s of this pool (1625-1647 vs 1638 frames/s, profiles/r1n_*)
    int sm_share[PAFUSE_MAX_PARTS] = {0};
    bool shares_fixed = false;       // set through pafuse_set_part_streams
    cudaStream_t side[PAFUSE_MAX_PARTS] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[PAFUSE_MAX_PARTS] = {nullptr};
    float* pred = nullptr;           // [S,F,num_kps,3]
    size_t pred_cap = 0;             // floats
    bool committed = false;
    bool debug_simt = false;
    bool debug_simt_attn = false;
    bool fuse_ln = true;             // LayerNorms fused into the proj / fc2 GEMM epilogues where the row fits one tile
    bool fuse_mlp = false;           // fc1 + GELU + fc2 (+ those LayerNorms) in one kernel for the same parts
    Profiler prof;
};

namespace {

struct ProfScope {
    pafuse_ctx* ctx;
    cudaStream_t st;
    size_t idx = 0;
    bool active;
    ProfScope(pafuse_ctx* c, int cat, double work, cudaStream_t s, double bytes = -1.0) : ctx(c), st(s), active(c->prof.on) {
        if (!active) return;
        ProfRec r;
        r.cat = cat;
        r.work = work;
        r.bytes = bytes >= 0.0 ? bytes : (cat == CAT_GEMM || cat == CAT_ATTN ? 0.0 : work);
        r.a = c->prof.get();
        r.b = c->prof.get();
        cudaEventRecord(r.a, st);
        idx = c->prof.recs.size();
        c->prof.recs.push_back(r);
    }
    ~ProfScope() {
        if (active) cudaEventRecord(ctx->prof.recs[idx].b, st);
    }
};

void add_slot(Part& p, const std::string& name, size_t numel, bool gemm = false, bool derived = false) {
    Slot s;
    s.off = p.total;
    s.numel = numel;
    s.gemm = gemm;
    s.derived = derived;
    p.slots[name] = s;
    p.total += (numel + 63) / 64 * 64;   // 256-byte alignment (TMA needs 16 B on the bf16 copies)
}

// tensor table of one MixSTE2 denoiser (common/mixste.py:141-210)
void build_part_table(Part& p, int C, int J, int F, int depth) {
    p.C = C;
    p.J = J;
    size_t c = (size_t)C;
    add_slot(p, "Spatial_patch_to_embedding.weight", c * 5);
    add_slot(p, "Spatial_patch_to_embedding.bias", c);
    add_slot(p, "Spatial_pos_embed", (size_t)J * c);
    add_slot(p, "Temporal_pos_embed", (size_t)F * c);
    add_slot(p, "time_mlp.1.weight", 2 * c * c);
    add_slot(p, "time_mlp.1.bias", 2 * c);
    add_slot(p, "time_mlp.3.weight", 2 * c * c);
    add_slot(p, "time_mlp.3.bias", c);
    add_slot(p, "Spatial_norm.weight", c);
    add_slot(p, "Spatial_norm.bias", c);
    add_slot(p, "Temporal_norm.weight", c);
    add_slot(p, "Temporal_norm.bias", c);
    add_slot(p, "head.0.weight", c);
    add_slot(p, "head.0.bias", c);
    add_slot(p, "head.1.weight", 3 * c);
    add_slot(p, "head.1.bias", 3);
    const char* stacks[2] = {"STEblocks", "TTEblocks"};
    for (int st = 0; st < 2; ++st)
        for (int i = 0; i < depth; ++i) {
            std::string b = std::string(stacks[st]) + "." + std::to_string(i) + ".";
            add_slot(p, b + "norm1.weight", c);
            add_slot(p, b + "norm1.bias", c);
            add_slot(p, b + "attn.qkv.weight", 3 * c * c, true);
            add_slot(p, b + "attn.qkv.bias", 3 * c);
            // plane-ordered, head-padded copies the production qkv GEMM reads (launch_pack_qkv)
            const size_t hds = (size_t)attn_head_store(C / 8);
            add_slot(p, b + "attn.qkv.weight_planes", 24 * hds * c, true, true);
            add_slot(p, b + "attn.qkv.bias_planes", 24 * hds, false, true);
            add_slot(p, b + "attn.proj.weight", c * c, true);
            add_slot(p, b + "attn.proj.bias", c);
            add_slot(p, b + "norm2.weight", c);
            add_slot(p, b + "norm2.bias", c);
            add_slot(p, b + "mlp.fc1.weight", 2 * c * c, true);
            add_slot(p, b + "mlp.fc1.bias", 2 * c);
            add_slot(p, b + "mlp.fc2.weight", 2 * c * c, true);
            add_slot(p, b + "mlp.fc2.bias", c);
        }
}

template <typename T>
int dev_alloc(T** p, size_t n) {
    PAFUSE_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
    return 0;
}

// captured passes hold workspace / prediction pointers and the switches of the context: drop them when any of it changes
void drop_graphs(pafuse_ctx* ctx) {
    for (auto& kv : ctx->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    ctx->graphs.clear();
}

int ensure_workspace(pafuse_ctx* ctx, Workspace& w, long long rows_x_c) {
    if (rows_x_c <= w.rows_x_c) return 0;
    drop_graphs(ctx);
    cudaFree(w.x); cudaFree(w.a_hi); cudaFree(w.a_lo); cudaFree(w.qkv); cudaFree(w.pl_hi); cudaFree(w.pl_lo);
    cudaFree(w.o_hi); cudaFree(w.o_lo); cudaFree(w.h_hi); cudaFree(w.h_lo);
    w = Workspace();
    size_t n = (size_t)rows_x_c;
    if (dev_alloc(&w.x, n)) return PAFUSE_E_CUDA;
    if (dev_alloc(&w.a_hi, n) || dev_alloc(&w.a_lo, n)) return PAFUSE_E_CUDA;
    if (ctx->debug_simt_attn && dev_alloc(&w.qkv, 3 * n)) return PAFUSE_E_CUDA;
    // planes: rows * 24 * hds halves with rows * C = n: 3 * hds / hd * n, and hds / hd <= 16/13 for hd >= 4... bound by 4 * n
    w.plane_halves = 4 * n + 4096;
    if (dev_alloc(&w.pl_hi, w.plane_halves) || dev_alloc(&w.pl_lo, w.plane_halves)) return PAFUSE_E_CUDA;
    if (dev_alloc(&w.o_hi, n) || dev_alloc(&w.o_lo, n)) return PAFUSE_E_CUDA;
    if (dev_alloc(&w.h_hi, 2 * n) || dev_alloc(&w.h_lo, 2 * n)) return PAFUSE_E_CUDA;
    w.rows_x_c = rows_x_c;
    return 0;
}

int ensure_pred(pafuse_ctx* ctx, size_t floats) {
    if (floats <= ctx->pred_cap) return 0;
    drop_graphs(ctx);
    cudaFree(ctx->pred);
    ctx->pred = nullptr;
    ctx->pred_cap = 0;
    if (dev_alloc(&ctx->pred, floats)) return PAFUSE_E_CUDA;
    ctx->pred_cap = floats;
    return 0;
}

// algorithmic DRAM bytes of one GEMM launch: fp16 hi/lo operands (4 B per element) read once, W once, and the
// epilogue's traffic (planes / hidden hi+lo: 4 B; residual read + write: 8 B; + LayerNorm output hi+lo: 4 B)
double gemm_bytes(const GemmArgs& g) {
    const double mn = (double)g.M * g.N;
    double out = 4.0 * mn;
    if (g.epilogue == EPI_RESID) out = 8.0 * mn;
    if (g.epilogue == EPI_RESID_LN) out = 12.0 * mn;
    return 4.0 * (double)g.M * g.K + 4.0 * (double)g.N * g.K + out;
}

int run_gemm(pafuse_ctx* ctx, const GemmArgs& g, cudaStream_t st) {
    ProfScope ps(ctx, CAT_GEMM, 2.0 * (double)g.M * g.N * g.K, st, gemm_bytes(g));
    return ctx->debug_simt ? launch_gemm_simt(g, st) : launch_gemm_tcgen05(g, st);
}

// One part denoiser (embed -> 2*depth blocks -> head) over the Sc sequences starting at s0, on stream st with
// workspace w; sm_limit > 0 sizes the persistent kernels for a share of the SMs.
int run_part(pafuse_ctx* ctx, int pi, Workspace& w, const float* x2d, const float* x2d_flip, const float* x3d,
             int apply_clamp, int R, int H, int s0, int Sc, float* pred, cudaStream_t st, int sm_limit) {
    const pafuse_config& cfg = ctx->cfg;
    const int F = cfg.frames;
    Part& p = ctx->parts[pi];
    const int C = p.C, J = p.J;
    const long long M = (long long)Sc * F * J;

    EmbedParams e;
    e.M = M; e.C = C; e.J = J; e.F = F; e.H = H; e.R = R; e.s0 = s0; e.num_kps = cfg.num_kps;
    e.apply_clamp = apply_clamp;
    e.clamp = (float)(1.1 * (double)cfg.scale);
    e.scale = cfg.scale;
    e.x2d = x2d; e.x2d_flip = x2d_flip; e.x3d = x3d;
    e.part_joints = p.joints_dev; e.flip_perm = ctx->flip_perm_dev;
    e.we = p.w("Spatial_patch_to_embedding.weight"); e.be = p.w("Spatial_patch_to_embedding.bias");
    e.spos = p.w("Spatial_pos_embed"); e.temb = p.temb; e.x = w.x;
    // norm1 of STE block 0 rides on the embedding kernel (the row is in registers there)
    e.g1 = p.w("STEblocks.0.norm1.weight"); e.b1 = p.w("STEblocks.0.norm1.bias"); e.eps1 = 1e-6f;
    e.out_hi = w.a_hi; e.out_lo = w.a_lo;
    {
        ProfScope ps(ctx, CAT_EMBED_HEAD, 8.0 * (double)M * C, st);
        if (int rc = launch_embed(e, st)) return rc;
    }

    // Parts whose rows fit one GEMM tile (C <= 256: face, hands) run proj and fc2 with the LayerNorms that
    // follow them fused into the epilogue (EPI_RESID_LN); the others keep the separate ln_chain launches.
    const bool fuse = ctx->fuse_ln && !ctx->debug_simt && gemm_can_fuse_ln(C);
    const int nblk = 2 * cfg.depth;
    for (int blk = 0; blk < nblk; ++blk) {
        const bool temporal = blk & 1;
        const std::string b = std::string(temporal ? "TTEblocks." : "STEblocks.") + std::to_string(blk / 2) + ".";
        LnParams l;
        l.M = M; l.C = C; l.J = J; l.F = F; l.x = w.x;
        l.g0 = l.b0 = nullptr; l.add_f = nullptr; l.eps0 = 1e-6f; l.eps1 = 1e-6f;
        l.out_hi = w.a_hi; l.out_lo = w.a_lo;
        if (!fuse && blk > 0) {
            // shared norm of the previous block (+ Temporal_pos_embed before TTE 0), then norm1 -> hi/lo
            if (blk > 0) {
                const char* sn = temporal ? "Spatial_norm" : "Temporal_norm";   // norm that closed the previous block
                l.g0 = p.w(std::string(sn) + ".weight");
                l.b0 = p.w(std::string(sn) + ".bias");
                if (blk == 1) l.add_f = p.w("Temporal_pos_embed");
            }
            l.g1 = p.w(b + "norm1.weight"); l.b1 = p.w(b + "norm1.bias");
            ProfScope ps(ctx, CAT_LN, (blk > 0 ? 12.0 : 8.0) * (double)M * C, st);
            if (int rc = launch_ln_chain(l, st)) return rc;
        }

        GemmArgs g;
        g.a_hi = w.a_hi; g.a_lo = w.a_lo;
        g.M = M; g.K = C; g.sm_limit = sm_limit;
        const double L = temporal ? (double)F : (double)J;
        if (ctx->debug_simt_attn) {
            // debug: fp32 qkv + CUDA-core attention
            if (!w.qkv) {
                set_last_error("debug attention must be selected before the first pass (workspace has no fp32 qkv)");
                return PAFUSE_E_STATE;
            }
            g.w_hi = p.wh(b + "attn.qkv.weight"); g.w_lo = p.wl(b + "attn.qkv.weight");
            g.bias = p.w(b + "attn.qkv.bias");
            g.out_f32 = w.qkv; g.out_hi = g.out_lo = nullptr;
            g.N = 3 * C; g.epilogue = EPI_F32;
            if (int rc = run_gemm(ctx, g, st)) return rc;
            AttnParams a;
            a.qkv = w.qkv; a.out_hi = w.o_hi; a.out_lo = w.o_lo;
            a.S = Sc; a.F = F; a.J = J; a.C = C; a.temporal = temporal ? 1 : 0; a.scale = 0.f;
            ProfScope ps(ctx, CAT_ATTN, 4.0 * (double)M * L * C, st);
            if (int rc = launch_attention(a, st)) return rc;
        } else {
            // the qkv GEMM writes the fp16 hi/lo head planes the attention kernel loads
            AttnPlanes pl;
            pl.hi = w.pl_hi; pl.lo = w.pl_lo; pl.rows_cap = M; pl.hds = attn_head_store(C / 8);
            if ((size_t)24 * M * pl.hds > w.plane_halves) {
                set_last_error("attention plane workspace too small");
                return PAFUSE_E_STATE;
            }
            g.w_hi = p.wh(b + "attn.qkv.weight_planes"); g.w_lo = p.wl(b + "attn.qkv.weight_planes");
            g.bias = p.w(b + "attn.qkv.bias_planes");
            g.out_f32 = nullptr; g.out_hi = g.out_lo = nullptr;
            g.N = 24 * pl.hds; g.epilogue = EPI_PLANES; g.planes = pl;
            {
                ProfScope ps(ctx, CAT_GEMM, 2.0 * (double)M * 3 * C * C, st, gemm_bytes(g));   // algorithmic N = 3C (pad columns not counted)
                if (int rc = ctx->debug_simt ? launch_gemm_simt(g, st) : launch_gemm_tcgen05(g, st)) return rc;
            }
            ProfScope ps(ctx, CAT_ATTN, 4.0 * (double)M * L * C, st, 4.0 * 24.0 * (double)M * pl.hds + 4.0 * (double)M * C);
            if (int rc = launch_attention_tc(pl, w.o_hi, w.o_lo, Sc, F, J, C, temporal ? 1 : 0, st, sm_limit)) return rc;
        }

        // proj: x += o W^T + b, then norm2 -> hi/lo
        g = GemmArgs();
        g.M = M; g.sm_limit = sm_limit;
        g.a_hi = w.o_hi; g.a_lo = w.o_lo;
        g.w_hi = p.wh(b + "attn.proj.weight"); g.w_lo = p.wl(b + "attn.proj.weight");
        g.bias = p.w(b + "attn.proj.bias");
        g.out_f32 = w.x; g.N = C; g.K = C; g.epilogue = EPI_RESID;
        if (fuse) {
            g.epilogue = EPI_RESID_LN;
            g.out_hi = w.a_hi; g.out_lo = w.a_lo;
            g.ln.x = w.x;
            g.ln.g1 = p.w(b + "norm2.weight"); g.ln.b1 = p.w(b + "norm2.bias"); g.ln.eps1 = 1e-6f;
            g.ln.J = J; g.ln.F = F;
        }
        if (int rc = run_gemm(ctx, g, st)) return rc;
        if (!fuse) {
            l.g0 = l.b0 = nullptr; l.add_f = nullptr;
            l.g1 = p.w(b + "norm2.weight"); l.b1 = p.w(b + "norm2.bias");
            ProfScope ps(ctx, CAT_LN, 8.0 * (double)M * C, st);
            if (int rc = launch_ln_chain(l, st)) return rc;
        }

        if (fuse && ctx->fuse_mlp && blk + 1 < nblk && mlp_can_fuse(C)) {
            // fc1 + GELU + fc2 + residual + the norms that follow in ONE kernel: h stays in tensor memory
            const bool tnext = (blk + 1) & 1;
            const std::string bn = std::string(tnext ? "TTEblocks." : "STEblocks.") + std::to_string((blk + 1) / 2) + ".";
            const char* sn = temporal ? "Temporal_norm" : "Spatial_norm";       // norm that closes this block
            MlpArgs m;
            m.a_hi = w.a_hi; m.a_lo = w.a_lo;
            m.w1_hi = p.wh(b + "mlp.fc1.weight"); m.w1_lo = p.wl(b + "mlp.fc1.weight"); m.b1 = p.w(b + "mlp.fc1.bias");
            m.w2_hi = p.wh(b + "mlp.fc2.weight"); m.w2_lo = p.wl(b + "mlp.fc2.weight"); m.b2 = p.w(b + "mlp.fc2.bias");
            m.x = w.x; m.out_hi = w.a_hi; m.out_lo = w.a_lo;
            m.M = M; m.C = C; m.sm_limit = sm_limit;
            m.ln.x = w.x;
            m.ln.g0 = p.w(std::string(sn) + ".weight"); m.ln.b0 = p.w(std::string(sn) + ".bias"); m.ln.eps0 = 1e-6f;
            if (blk == 0) m.ln.add_f = p.w("Temporal_pos_embed");
            m.ln.g1 = p.w(bn + "norm1.weight"); m.ln.b1 = p.w(bn + "norm1.bias"); m.ln.eps1 = 1e-6f;
            m.ln.J = J; m.ln.F = F;
            ProfScope ps(ctx, CAT_GEMM, 2.0 * 2.0 * (double)M * C * 2 * C, st, 16.0 * (double)M * C + 16.0 * (double)C * C);
            if (int rc = launch_mlp_fused(m, st)) return rc;
            continue;
        }

        g = GemmArgs();
        g.M = M; g.sm_limit = sm_limit;
        g.a_hi = w.a_hi; g.a_lo = w.a_lo;
        g.w_hi = p.wh(b + "mlp.fc1.weight"); g.w_lo = p.wl(b + "mlp.fc1.weight");
        g.bias = p.w(b + "mlp.fc1.bias");
        g.out_f32 = nullptr; g.out_hi = w.h_hi; g.out_lo = w.h_lo;
        g.N = 2 * C; g.K = C; g.epilogue = EPI_GELU_SPLIT;
        if (int rc = run_gemm(ctx, g, st)) return rc;

        // fc2: x += h W^T + b; fused: the norm that closes this block (+ Temporal_pos_embed after STE 0) and
        // norm1 of the next block
        g = GemmArgs();
        g.M = M; g.sm_limit = sm_limit;
        g.a_hi = w.h_hi; g.a_lo = w.h_lo;
        g.w_hi = p.wh(b + "mlp.fc2.weight"); g.w_lo = p.wl(b + "mlp.fc2.weight");
        g.bias = p.w(b + "mlp.fc2.bias");
        g.out_f32 = w.x; g.out_hi = g.out_lo = nullptr;
        g.N = C; g.K = 2 * C; g.epilogue = EPI_RESID;
        if (fuse && blk + 1 < nblk) {
            const bool tnext = (blk + 1) & 1;
            const std::string bn = std::string(tnext ? "TTEblocks." : "STEblocks.") + std::to_string((blk + 1) / 2) + ".";
            const char* sn = temporal ? "Temporal_norm" : "Spatial_norm";       // norm that closes this block
            g.epilogue = EPI_RESID_LN;
            g.out_hi = w.a_hi; g.out_lo = w.a_lo;
            g.ln.x = w.x;
            g.ln.g0 = p.w(std::string(sn) + ".weight"); g.ln.b0 = p.w(std::string(sn) + ".bias"); g.ln.eps0 = 1e-6f;
            if (blk == 0) g.ln.add_f = p.w("Temporal_pos_embed");
            g.ln.g1 = p.w(bn + "norm1.weight"); g.ln.b1 = p.w(bn + "norm1.bias"); g.ln.eps1 = 1e-6f;
            g.ln.J = J; g.ln.F = F;
        }
        if (int rc = run_gemm(ctx, g, st)) return rc;
    }

    HeadParams h;
    h.M = M; h.C = C; h.J = J; h.F = F; h.s0 = s0; h.num_kps = cfg.num_kps; h.x = w.x;
    h.g0 = p.w("Temporal_norm.weight"); h.b0 = p.w("Temporal_norm.bias"); h.eps0 = 1e-6f;
    h.g1 = p.w("head.0.weight"); h.b1 = p.w("head.0.bias"); h.eps1 = 1e-5f;
    h.wh = p.w("head.1.weight"); h.bh = p.w("head.1.bias");
    h.part_joints = p.joints_dev; h.pred = pred;
    {
        ProfScope ps(ctx, CAT_EMBED_HEAD, 4.0 * (double)M * C, st);
        if (int rc = launch_head(h, st)) return rc;
    }
    return 0;
}

// SM shares of the side-by-side mode: proportional to J*C (the parts' DRAM bytes per sequence, which is what their
// time follows: body 26 %, face 43 %, hands 31 % of a pass), in whole CTA pairs.  PAFUSE_SM_SHARES="a,b,c" overrides.
void pick_sm_shares(pafuse_ctx* ctx, int num_sms) {
    const int n = ctx->num_parts;
    if (const char* e = getenv("PAFUSE_SM_SHARES")) {
        int v[PAFUSE_MAX_PARTS] = {0}, k = 0;
        const char* q = e;
        while (*q && k < n) {
            v[k++] = atoi(q);
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (k == n) {
            for (int i = 0; i < n; ++i) ctx->sm_share[i] = v[i] < 2 ? 2 : v[i] / 2 * 2;
            return;
        }
    }
    double tot = 0;
    for (int i = 0; i < n; ++i) tot += (double)ctx->parts[i].J * ctx->parts[i].C;
    int used = 0;
    for (int i = 0; i < n; ++i) {
        int v = (int)((double)ctx->parts[i].J * ctx->parts[i].C / tot * num_sms + 0.5) / 2 * 2;
        if (v < 2) v = 2;
        ctx->sm_share[i] = v;
        used += v;
    }
    // the rounding remainder goes to (or comes from) the largest share
    int big = 0;
    for (int i = 1; i < n; ++i)
        if (ctx->sm_share[i] > ctx->sm_share[big]) big = i;
    ctx->sm_share[big] += (num_sms - used) / 2 * 2;
}

// One pass of the three part denoisers over S_total sequences (R originals followed,
// when S_total == 2R, by their flip-TTA twins), chunked by cfg.max_seqs.
int run_denoisers(pafuse_ctx* ctx, const float* x2d, const float* x2d_flip, const float* x3d, int apply_clamp, int R,
                  int H, int S_total, const float* sinus, float* pred, cudaStream_t st) {
    const pafuse_config& cfg = ctx->cfg;
    const int F = cfg.frames;
    if (!ctx->committed) {
        set_last_error("weights not committed (call pafuse_commit_weights)");
        return PAFUSE_E_STATE;
    }
    // time embedding per part (identical for every row: the sampler uses one t per call)
    {
        int off = 0;
        for (int pi = 0; pi < ctx->num_parts; ++pi) {
            Part& p = ctx->parts[pi];
            ProfScope ps(ctx, CAT_EMBED_HEAD, 4.0 * 4.0 * p.C * p.C, st);
            if (int rc = launch_time_mlp(sinus + off, p.w("time_mlp.1.weight"), p.w("time_mlp.1.bias"),
                                         p.w("time_mlp.3.weight"), p.w("time_mlp.3.bias"), p.temb, p.C, st))
                return rc;
            off += p.C;
        }
    }
    int max_seqs = cfg.max_seqs > 0 ? cfg.max_seqs : 640;
    int chunk = S_total < max_seqs ? S_total : max_seqs;
    // side by side needs the production kernels (the persistent ones can be sized to a share) and no per-launch
    // timing (bench.py's profiled pass measures every kernel alone on the whole GPU)
    const bool side_by_side = ctx->part_streams && ctx->num_parts > 1 && !ctx->prof.on && !ctx->debug_simt &&
                              !ctx->debug_simt_attn;
    if (!side_by_side) {
        long long need = 0;
        for (int pi = 0; pi < ctx->num_parts; ++pi) {
            long long v = (long long)chunk * F * ctx->parts[pi].J * ctx->parts[pi].C;
            if (v > need) need = v;
        }
        if (int rc = ensure_workspace(ctx, ctx->ws[0], need)) return rc;
        for (int s0 = 0; s0 < S_total; s0 += chunk) {
            const int Sc = (S_total - s0) < chunk ? (S_total - s0) : chunk;
            for (int pi = 0; pi < ctx->num_parts; ++pi)
                if (int rc = run_part(ctx, pi, ctx->ws[0], x2d, x2d_flip, x3d, apply_clamp, R, H, s0, Sc, pred, st, 0))
                    return rc;
        }
        return 0;
    }

    if (!ctx->ev_fork) {
        int num_sms = 0;
        PAFUSE_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (!ctx->shares_fixed) pick_sm_shares(ctx, num_sms);
        PAFUSE_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (int pi = 1; pi < ctx->num_parts; ++pi) {
            PAFUSE_CUDA_OK(cudaStreamCreateWithFlags(&ctx->side[pi], cudaStreamNonBlocking));
            PAFUSE_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_join[pi], cudaEventDisableTiming));
        }
    }
    for (int pi = 0; pi < ctx->num_parts; ++pi)
        if (int rc = ensure_workspace(ctx, ctx->ws[pi], (long long)chunk * F * ctx->parts[pi].J * ctx->parts[pi].C))
            return rc;
    // fork: part 0 stays on the caller's stream, the others run on side streams that start after everything the
    // caller enqueued so far (inputs, time embeddings, the previous DDIM update) and are joined before returning
    PAFUSE_CUDA_OK(cudaEventRecord(ctx->ev_fork, st));
    for (int pi = 1; pi < ctx->num_parts; ++pi) PAFUSE_CUDA_OK(cudaStreamWaitEvent(ctx->side[pi], ctx->ev_fork, 0));
    ++g_pdl_suspend;
    int rc = 0;
    for (int s0 = 0; s0 < S_total && !rc; s0 += chunk) {
        const int Sc = (S_total - s0) < chunk ? (S_total - s0) : chunk;
        for (int pi = 0; pi < ctx->num_parts && !rc; ++pi)
            rc = run_part(ctx, pi, ctx->ws[pi], x2d, x2d_flip, x3d, apply_clamp, R, H, s0, Sc, pred,
                          pi == 0 ? st : ctx->side[pi], ctx->sm_share[pi]);
    }
    --g_pdl_suspend;
    for (int pi = 1; pi < ctx->num_parts; ++pi) {                      // join even after an error: no dangling work
        cudaEventRecord(ctx->ev_join[pi], ctx->side[pi]);
        cudaStreamWaitEvent(st, ctx->ev_join[pi], 0);
    }
    return rc;
}

// run_denoisers, replayed from a CUDA graph when the pass is small enough to be launch-bound.  First call with a key:
// direct (it may allocate workspaces and set function attributes, neither of which can be captured); second call:
// captured on the context's private stream, instantiated and launched; afterwards: one cudaGraphLaunch per pass.
int run_denoisers_cached(pafuse_ctx* ctx, const float* x2d, const float* x2d_flip, const float* x3d, int apply_clamp, int R,
                         int H, int S_total, const float* sinus, float* pred, cudaStream_t st) {
    const bool eligible = ctx->graph_max_seqs > 0 && S_total <= ctx->graph_max_seqs && !ctx->prof.on && !ctx->part_streams &&
                          ctx->committed;
    if (!eligible) return run_denoisers(ctx, x2d, x2d_flip, x3d, apply_clamp, R, H, S_total, sinus, pred, st);
    if (ctx->graph_epoch != g_switch_epoch) {
        drop_graphs(ctx);
        ctx->graph_epoch = g_switch_epoch;
    }
    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.x2d = x2d; key.x2d_flip = x2d_flip; key.x3d = x3d; key.sinus = sinus; key.pred = pred;
    key.apply_clamp = apply_clamp; key.R = R; key.H = H; key.S_total = S_total;
    if (ctx->graphs.size() > 64 && ctx->graphs.find(key) == ctx->graphs.end()) drop_graphs(ctx);   // bound the cache
    GraphEntry& e = ctx->graphs[key];
    if (e.exec) {
        PAFUSE_CUDA_OK(cudaGraphLaunch(e.exec, st));
        g_launch_count += e.launches;
        ++ctx->graph_replays;
        return 0;
    }
    if (e.seen++ == 0) return run_denoisers(ctx, x2d, x2d_flip, x3d, apply_clamp, R, H, S_total, sinus, pred, st);
    if (!ctx->cap_stream) PAFUSE_CUDA_OK(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
    const long long before = g_launch_count;
    PAFUSE_CUDA_OK(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_denoisers(ctx, x2d, x2d_flip, x3d, apply_clamp, R, H, S_total, sinus, pred, ctx->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
    const long long captured = g_launch_count - before;
    g_launch_count = before;                                    // nothing ran yet
    if (rc != 0 || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->graphs.erase(key);
        if (rc != 0) return rc;
        // capture refused (e.g. a first-use attribute call slipped in): run directly, try again next time
        return run_denoisers(ctx, x2d, x2d_flip, x3d, apply_clamp, R, H, S_total, sinus, pred, st);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || !exec) {
        cudaGetLastError();
        ctx->graphs.erase(key);
        return run_denoisers(ctx, x2d, x2d_flip, x3d, apply_clamp, R, H, S_total, sinus, pred, st);
    }
    e.exec = exec;
    e.launches = captured;
    PAFUSE_CUDA_OK(cudaGraphLaunch(exec, st));
    g_launch_count += captured;
    ++ctx->graph_replays;
    return 0;
}

// resident copy of a small host table: uploaded only when its contents change (the upload reads pageable memory and
// therefore drains the stream first -- once per table, not once per call)
int sync_table(std::vector<int>& host, int* dev, const int* src, int n, cudaStream_t st) {
    if ((int)host.size() == n && memcmp(host.data(), src, (size_t)n * sizeof(int)) == 0) return 0;
    host.assign(src, src + n);
    PAFUSE_CUDA_OK(cudaMemcpyAsync(dev, host.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    PAFUSE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

bool check_ctx(pafuse_ctx* ctx) {
    if (!ctx) {
        set_last_error("null context");
        return false;
    }
    return true;
}

}  // namespace

extern "C" {

const char* pafuse_last_error(void) { return g_err; }
const char* pafuse_version(void) { return "pafuse_b200 0.2 (sm_100a, tcgen05 f16x3, cta_group::2)"; }
int64_t pafuse_launch_count(void) { return (int64_t)g_launch_count; }

int pafuse_create(const pafuse_config* cfg, pafuse_ctx** out) {
    if (!cfg || !out) {
        set_last_error("pafuse_create: null argument");
        return PAFUSE_E_ARG;
    }
    if (cfg->num_parts < 1 || cfg->num_parts > PAFUSE_MAX_PARTS || cfg->heads != 8 || cfg->frames < 1 ||
        cfg->depth < 1 || cfg->num_kps < 1 || !cfg->flip_perm) {
        set_last_error("pafuse_create: bad config (parts=%d heads=%d frames=%d depth=%d)", cfg->num_parts, cfg->heads,
                       cfg->frames, cfg->depth);
        return PAFUSE_E_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_last_error("pafuse_create: no CUDA device (this library has no CPU fallback)");
        return PAFUSE_E_CUDA;
    }
    pafuse_ctx* ctx = new pafuse_ctx();
    ctx->cfg = *cfg;
    ctx->num_parts = cfg->num_parts;
    PAFUSE_CUDA_OK(cudaGetDevice(&ctx->device));
    cudaDeviceProp prop;
    PAFUSE_CUDA_OK(cudaGetDeviceProperties(&prop, ctx->device));
    if (prop.major != 10) {
        set_last_error("pafuse_create: device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor);
        delete ctx;
        return PAFUSE_E_CUDA;
    }
    for (int pi = 0; pi < cfg->num_parts; ++pi) {
        Part& p = ctx->parts[pi];
        int C = cfg->part_channels[pi], J = cfg->part_num_joints[pi];
        if (C % 32 != 0 || C > 384 || J < 1 || !cfg->part_joints[pi]) {
            set_last_error("pafuse_create: part %d has unsupported C=%d J=%d", pi, C, J);
            delete ctx;
            return PAFUSE_E_ARG;
        }
        build_part_table(p, C, J, cfg->frames, cfg->depth);
        if (dev_alloc(&p.f32, p.total) || dev_alloc(&p.hi, p.total) || dev_alloc(&p.lo, p.total) ||
            dev_alloc(&p.joints_dev, (size_t)J) || dev_alloc(&p.temb, (size_t)C))
            return PAFUSE_E_CUDA;
        PAFUSE_CUDA_OK(cudaMemset(p.f32, 0, p.total * sizeof(float)));
        PAFUSE_CUDA_OK(cudaMemcpy(p.joints_dev, cfg->part_joints[pi], J * sizeof(int), cudaMemcpyHostToDevice));
        ctx->cfg.part_joints[pi] = nullptr;   // host pointer not retained
    }
    if (dev_alloc(&ctx->flip_perm_dev, (size_t)cfg->num_kps) || dev_alloc(&ctx->conn_dev, (size_t)cfg->num_kps) ||
        dev_alloc(&ctx->conn_rows_dev, (size_t)cfg->num_kps) || dev_alloc(&ctx->part_of_joint_dev, (size_t)cfg->num_kps) ||
        dev_alloc(&ctx->root_of_joint_dev, (size_t)cfg->num_kps) || dev_alloc(&ctx->absmax_dev, (size_t)1))
        return PAFUSE_E_CUDA;
    PAFUSE_CUDA_OK(cudaMemcpy(ctx->flip_perm_dev, cfg->flip_perm, cfg->num_kps * sizeof(int), cudaMemcpyHostToDevice));
    ctx->cfg.flip_perm = nullptr;
    if (int rc = gemm_init()) return rc;
    if (const char* e = getenv("PAFUSE_FUSE_LN")) ctx->fuse_ln = atoi(e) != 0;
    if (const char* e = getenv("PAFUSE_PART_STREAMS")) ctx->part_streams = atoi(e) != 0;
    if (const char* e = getenv("PAFUSE_FUSE_MLP")) ctx->fuse_mlp = atoi(e) != 0;
    if (const char* e = getenv("PAFUSE_GRAPH_MAX_SEQS")) ctx->graph_max_seqs = atoi(e);
    *out = ctx;
    return 0;
}

void pafuse_destroy(pafuse_ctx* ctx) {
    if (!ctx) return;
    for (int pi = 0; pi < ctx->num_parts; ++pi) {
        Part& p = ctx->parts[pi];
        cudaFree(p.f32); cudaFree(p.hi); cudaFree(p.lo); cudaFree(p.joints_dev); cudaFree(p.temb);
    }
    for (Workspace& w : ctx->ws) {
        cudaFree(w.x); cudaFree(w.a_hi); cudaFree(w.a_lo); cudaFree(w.qkv); cudaFree(w.pl_hi); cudaFree(w.pl_lo);
        cudaFree(w.o_hi); cudaFree(w.o_lo); cudaFree(w.h_hi); cudaFree(w.h_lo);
    }
    drop_graphs(ctx);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    cudaFree(ctx->part_of_joint_dev); cudaFree(ctx->root_of_joint_dev); cudaFree(ctx->absmax_dev);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int pi = 0; pi < PAFUSE_MAX_PARTS; ++pi) {
        if (ctx->side[pi]) cudaStreamDestroy(ctx->side[pi]);
        if (ctx->ev_join[pi]) cudaEventDestroy(ctx->ev_join[pi]);
    }
    cudaFree(ctx->flip_perm_dev); cudaFree(ctx->conn_dev); cudaFree(ctx->conn_rows_dev); cudaFree(ctx->pred);
    delete ctx;
}

int pafuse_set_weight(pafuse_ctx* ctx, int32_t part, const char* name, const float* data, int64_t numel,
                      int32_t on_device) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (part < 0 || part >= ctx->num_parts || !name || !data) {
        set_last_error("pafuse_set_weight: bad argument (part=%d)", part);
        return PAFUSE_E_ARG;
    }
    Part& p = ctx->parts[part];
    auto it = p.slots.find(name);
    if (it == p.slots.end() || it->second.derived) {
        set_last_error("pafuse_set_weight: unknown tensor '%s'", name);
        return PAFUSE_E_ARG;
    }
    if ((size_t)numel != it->second.numel) {
        set_last_error("pafuse_set_weight: '%s' expects %zu elements, got %lld", name, it->second.numel, (long long)numel);
        return PAFUSE_E_ARG;
    }
    PAFUSE_CUDA_OK(cudaMemcpy(p.f32 + it->second.off, data, (size_t)numel * sizeof(float),
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    it->second.set = true;
    ctx->committed = false;
    return 0;
}

int pafuse_commit_weights(pafuse_ctx* ctx, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    for (int pi = 0; pi < ctx->num_parts; ++pi) {
        Part& p = ctx->parts[pi];
        for (auto& kv : p.slots) {
            if (!kv.second.set && !kv.second.derived) {
                set_last_error("pafuse_commit_weights: part %d tensor '%s' was never set", pi, kv.first.c_str());
                return PAFUSE_E_STATE;
            }
        }
        const char* stacks[2] = {"STEblocks.", "TTEblocks."};
        for (int sk = 0; sk < 2; ++sk)
            for (int i = 0; i < ctx->cfg.depth; ++i) {
                const std::string b = std::string(stacks[sk]) + std::to_string(i) + ".attn.qkv.";
                if (int rc = launch_pack_qkv(p.f32 + p.slots.at(b + "weight").off, p.f32 + p.slots.at(b + "bias").off,
                                             p.f32 + p.slots.at(b + "weight_planes").off,
                                             p.f32 + p.slots.at(b + "bias_planes").off, p.C, p.C / 8,
                                             attn_head_store(p.C / 8), st))
                    return rc;
            }
        // fp16 range of the pre-scaled operands: |w| * WEIGHT_SCALE must stay below 65504, else the hi half would
        // saturate silently (the fp32 reference has no such limit).  NaN weights pass: they propagate as NaN.
        PAFUSE_CUDA_OK(cudaMemsetAsync(ctx->absmax_dev, 0, sizeof(unsigned int), st));
        for (auto& kv : p.slots) {
            if (kv.second.gemm) {
                if (int rc = launch_absmax(p.f32 + kv.second.off, kv.second.numel, ctx->absmax_dev, st)) return rc;
                if (int rc = launch_split_weights(p.f32 + kv.second.off, p.hi + kv.second.off, p.lo + kv.second.off,
                                                  kv.second.numel, st))
                    return rc;
            }
        }
        unsigned int bits = 0;
        PAFUSE_CUDA_OK(cudaMemcpyAsync(&bits, ctx->absmax_dev, sizeof(bits), cudaMemcpyDeviceToHost, st));
        PAFUSE_CUDA_OK(cudaStreamSynchronize(st));
        float amax;
        memcpy(&amax, &bits, sizeof(amax));
        if (amax == amax && amax * WEIGHT_SCALE > 65504.0f) {
            set_last_error("pafuse_commit_weights: part %d has a GEMM weight of magnitude %g; the f16x3 operand format holds "
                           "|w| < %g", pi, (double)amax, (double)(65504.0f / WEIGHT_SCALE));
            return PAFUSE_E_ARG;
        }
    }
    drop_graphs(ctx);
    ctx->committed = true;
    return 0;
}

int pafuse_pred_parts(pafuse_ctx* ctx, const float* x2d, const float* x3d, const float* sinus, float* out, int32_t B,
                      int32_t H, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (B == 0) return 0;
    if (!x2d || !x3d || !sinus || !out || B < 0 || H < 1) {
        set_last_error("pafuse_pred_parts: bad argument");
        return PAFUSE_E_ARG;
    }
    if (B == 0) return 0;
    return run_denoisers_cached(ctx, x2d, nullptr, x3d, 0, B * H, H, B * H, sinus, out, (cudaStream_t)stream);
}

int pafuse_ddim_step(pafuse_ctx* ctx, const float* x2d, const float* x2d_flip, const float* sinus, float* img,
                     const float* noise, float* x0_out, int64_t x0_batch_stride, int32_t B, int32_t H, int32_t flip,
                     int32_t last, double sqrt_recip, double sqrt_recipm1, double c64, float sqrt_an, float c,
                     float sigma, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (B == 0) return 0;
    if (!x2d || !sinus || !img || !x0_out || B < 0 || H < 1 || (flip && !x2d_flip) || (!last && !noise)) {
        set_last_error("pafuse_ddim_step: bad argument");
        return PAFUSE_E_ARG;
    }
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const pafuse_config& cfg = ctx->cfg;
    const int R = B * H;
    const int S_total = flip ? 2 * R : R;
    if (int rc = ensure_pred(ctx, (size_t)S_total * cfg.frames * cfg.num_kps * 3)) return rc;
    if (int rc = run_denoisers_cached(ctx, x2d, x2d_flip, img, 1, R, H, S_total, sinus, ctx->pred, st)) return rc;
    DdimParams d;
    d.R = R; d.F = cfg.frames; d.H = H; d.num_kps = cfg.num_kps; d.flip = flip ? 1 : 0; d.last = last ? 1 : 0;
    d.pred = ctx->pred; d.flip_perm = ctx->flip_perm_dev; d.img = img; d.noise = noise;
    d.x0_out = x0_out; d.x0_batch_stride = x0_batch_stride;
    d.scale = cfg.scale; d.clamp = (float)(1.1 * (double)cfg.scale);
    d.sqrt_recip = sqrt_recip; d.sqrt_recipm1 = sqrt_recipm1; d.c64 = c64;
    d.sqrt_an = sqrt_an; d.c = c; d.sigma = sigma;
    ProfScope ps(ctx, CAT_DDIM, (flip ? 24.0 : 20.0) * (double)R * cfg.frames * cfg.num_kps * 3, st);
    return launch_ddim_step(d, st);
}

int pafuse_wb_pose_from_parts(pafuse_ctx* ctx, float* pose, float* out, const int32_t* conn_of_joint, int64_t poses,
                              int32_t mutate_input, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!pose || !out || !conn_of_joint || pose == out || poses < 0) {
        set_last_error("pafuse_wb_pose_from_parts: bad argument (in-place use is not supported)");
        return PAFUSE_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nk = ctx->cfg.num_kps;
    if (poses == 0) return 0;
    std::vector<int> rows;
    for (int g = 0; g < nk; ++g) {
        int r = conn_of_joint[g];
        if (r >= nk) {
            set_last_error("pafuse_wb_pose_from_parts: connection index %d out of range", r);
            return PAFUSE_E_ARG;
        }
        bool seen = false;
        for (int v : rows) seen |= (v == r);
        if (r >= 0 && !seen) rows.push_back(r);
    }
    // the tables are constant per dataset: resident on the device, re-uploaded (with the one stream drain that reading
    // pageable memory costs) only when the caller passes a different table
    if (int rc = sync_table(ctx->conn_host, ctx->conn_dev, conn_of_joint, nk, st)) return rc;
    {
        ProfScope ps(ctx, CAT_POST, 8.0 * (double)poses * nk * 3, st);
        if (int rc = launch_reassemble(pose, out, ctx->conn_dev, poses, nk, st)) return rc;
    }
    if (mutate_input && !rows.empty()) {
        if (int rc = sync_table(ctx->conn_rows_host, ctx->conn_rows_dev, rows.data(), (int)rows.size(), st)) return rc;
        if (int rc = launch_negate_rows(pose, ctx->conn_rows_dev, (int)rows.size(), poses, nk, st)) return rc;
    }
    return 0;
}

int pafuse_project_to_2d(pafuse_ctx* ctx, const float* X, const float* cam, float* out, int64_t n_cams,
                         int64_t pts_per_cam, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!X || !cam || !out || n_cams < 0 || pts_per_cam < 1) {
        set_last_error("pafuse_project_to_2d: bad argument");
        return PAFUSE_E_ARG;
    }
    return launch_project(X, cam, out, n_cams * pts_per_cam, pts_per_cam, (cudaStream_t)stream);
}

int pafuse_aggregate(pafuse_ctx* ctx, const float* pred, const float* traj, const float* cam, int32_t cam_per_clip,
                     const float* x2d, float* jagg, float* pagg, int32_t* select, float* reproj, int32_t B, int32_t K,
                     int32_t H, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!pred || !cam || !x2d || !jagg || !pagg || B < 0 || K < 1 || H < 1) {
        set_last_error("pafuse_aggregate: bad argument");
        return PAFUSE_E_ARG;
    }
    AggParams a;
    a.B = B; a.K = K; a.H = H; a.F = ctx->cfg.frames; a.J = ctx->cfg.num_kps; a.cam_per_clip = cam_per_clip;
    a.pred = pred; a.traj = traj; a.cam = cam; a.x2d = x2d; a.jagg = jagg; a.pagg = pagg; a.select = select;
    a.reproj = reproj;
    const double per_pose = 12.0 * a.F * a.J;
    ProfScope ps(ctx, CAT_POST, per_pose * ((double)B * K * H + 2.0 * B * K), (cudaStream_t)stream);
    return launch_aggregate(a, (cudaStream_t)stream);
}

int pafuse_mpjpe_metrics(pafuse_ctx* ctx, const float* pred, const float* target, const float* traj, const float* cam,
                         int32_t cam_per_clip, const float* x2d, const float* reproj, double* sums, int32_t B, int32_t K,
                         int32_t H, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!pred || !target || !sums || (x2d && !reproj && !cam) || B < 0 || K < 1 || H < 1 || H > 256 || (long long)B * K > 65535) {
        set_last_error("pafuse_mpjpe_metrics: bad argument (B*K <= 65535, H <= 256; a 2D target needs reproj or cam)");
        return PAFUSE_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PAFUSE_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)K * (3 + H) * sizeof(double), st));
    if (B == 0) return 0;
    AggParams a;
    a.B = B; a.K = K; a.H = H; a.F = ctx->cfg.frames; a.J = ctx->cfg.num_kps; a.cam_per_clip = cam_per_clip;
    a.pred = pred; a.traj = traj; a.cam = cam; a.x2d = x2d; a.jagg = nullptr; a.pagg = nullptr; a.select = nullptr;
    a.reproj = nullptr;
    ProfScope ps(ctx, CAT_POST, 12.0 * a.F * a.J * ((double)B * K * H + B), st);
    return launch_metrics(a, target, reproj, sums, st);
}

int pafuse_mpjpe_metrics_parts(pafuse_ctx* ctx, const float* pred, const float* target, const int32_t* part_of_joint,
                               const int32_t* root_of_joint, int32_t n_parts, double* sums, int32_t B, int32_t K, int32_t H,
                               void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    const int nk = ctx->cfg.num_kps;
    if (!pred || !target || !part_of_joint || !root_of_joint || !sums || n_parts < 1 || n_parts > 16 || B < 0 || K < 1 ||
        H < 1 || H > 256 || (long long)B * K > 65535) {
        set_last_error("pafuse_mpjpe_metrics_parts: bad argument (1 <= n_parts <= 16, B*K <= 65535, H <= 256)");
        return PAFUSE_E_ARG;
    }
    for (int j = 0; j < nk; ++j)
        if (part_of_joint[j] >= n_parts || (part_of_joint[j] >= 0 && (root_of_joint[j] < 0 || root_of_joint[j] >= nk))) {
            set_last_error("pafuse_mpjpe_metrics_parts: joint %d has part %d / root %d out of range", j, part_of_joint[j],
                           root_of_joint[j]);
            return PAFUSE_E_ARG;
        }
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = sync_table(ctx->part_of_joint_host, ctx->part_of_joint_dev, part_of_joint, nk, st)) return rc;
    if (int rc = sync_table(ctx->root_of_joint_host, ctx->root_of_joint_dev, root_of_joint, nk, st)) return rc;
    PAFUSE_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)K * (H + 1) * n_parts * sizeof(double), st));
    if (B == 0) return 0;
    AggParams a;
    a.B = B; a.K = K; a.H = H; a.F = ctx->cfg.frames; a.J = nk; a.cam_per_clip = 0;
    a.pred = pred; a.traj = nullptr; a.cam = nullptr; a.x2d = nullptr; a.jagg = nullptr; a.pagg = nullptr; a.select = nullptr;
    a.reproj = nullptr;
    ProfScope ps(ctx, CAT_POST, 12.0 * a.F * a.J * ((double)B * K * H + B), st);
    return launch_metrics_parts(a, target, ctx->part_of_joint_dev, ctx->root_of_joint_dev, n_parts, sums, st);
}

int pafuse_randn(pafuse_ctx* ctx, uint64_t seed, uint64_t draw, int64_t base, float* out, int64_t rows, int64_t row_len,
                 int64_t row_stride, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (rows == 0 || row_len == 0) return 0;
    if (!out || rows < 0 || row_len < 0 || base < 0 || row_stride < row_len) {
        set_last_error("pafuse_randn: bad argument (rows=%lld row_len=%lld row_stride=%lld base=%lld)", (long long)rows,
                       (long long)row_len, (long long)row_stride, (long long)base);
        return PAFUSE_E_ARG;
    }
    ProfScope ps(ctx, CAT_DDIM, 4.0 * (double)rows * row_len, (cudaStream_t)stream);
    return launch_randn_philox(out, seed, draw, base, rows, row_len, row_stride, (cudaStream_t)stream);
}

int pafuse_set_graph_max_seqs(pafuse_ctx* ctx, int32_t max_seqs) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    ctx->graph_max_seqs = max_seqs < 0 ? 0 : max_seqs;
    drop_graphs(ctx);
    return 0;
}

int64_t pafuse_graph_replays(pafuse_ctx* ctx) { return ctx ? (int64_t)ctx->graph_replays : 0; }

int pafuse_prepare_clips(pafuse_ctx* ctx, const float* seq, int64_t T, float* clips, float* clips_flip, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (T == 0) return 0;
    if (!seq || !clips || T < 0) {
        set_last_error("pafuse_prepare_clips: bad argument");
        return PAFUSE_E_ARG;
    }
    const int F = ctx->cfg.frames;
    const long long n = (T + F - 1) / F;
    ProfScope ps(ctx, CAT_POST, (clips_flip ? 24.0 : 16.0) * (double)n * F * ctx->cfg.num_kps, (cudaStream_t)stream);
    return launch_prepare_clips(seq, T, F, ctx->cfg.num_kps, ctx->flip_perm_dev, clips, clips_flip, n, (cudaStream_t)stream);
}

int pafuse_stitch_clips(pafuse_ctx* ctx, const float* pred, int64_t n_clips, int32_t K, int32_t H, int64_t T, float* out,
                        void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (T == 0) return 0;
    const int F = ctx->cfg.frames;
    if (!pred || !out || K < 1 || H < 1 || T < 0 || n_clips != (T + F - 1) / F) {
        set_last_error("pafuse_stitch_clips: bad argument (n_clips=%lld must be ceil(T/F) for T=%lld)", (long long)n_clips,
                       (long long)T);
        return PAFUSE_E_ARG;
    }
    ProfScope ps(ctx, CAT_POST, 24.0 * (double)K * H * T * ctx->cfg.num_kps, (cudaStream_t)stream);
    return launch_stitch_clips(pred, out, n_clips, K, H, F, ctx->cfg.num_kps, T, (cudaStream_t)stream);
}

int pafuse_keypoints_from_detections(pafuse_ctx* ctx, const float* raw, int64_t T, int32_t width, int32_t height, float* kp,
                                     void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (T == 0) return 0;
    if (!raw || !kp || T < 0 || width < 1 || height < 1 || ctx->cfg.num_kps < 14) {
        set_last_error("pafuse_keypoints_from_detections: bad argument");
        return PAFUSE_E_ARG;
    }
    return launch_keypoints(raw, kp, T, ctx->cfg.num_kps, width, height, (cudaStream_t)stream);
}

int pafuse_set_debug_simt_gemm(pafuse_ctx* ctx, int32_t enable) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    ctx->debug_simt = enable != 0;
    drop_graphs(ctx);
    return 0;
}

int pafuse_set_debug_simt_attention(pafuse_ctx* ctx, int32_t enable) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    ctx->debug_simt_attn = enable != 0;
    drop_graphs(ctx);
    ctx->ws[0].rows_x_c = 0;     // the fp32 qkv buffer exists only in debug mode: force a re-allocation
    return 0;
}

int pafuse_set_gemm_cta_group(int32_t cta_group) {
    if (cta_group != 1 && cta_group != 2) {
        set_last_error("pafuse_set_gemm_cta_group: cta_group must be 1 or 2");
        return PAFUSE_E_ARG;
    }
    if (int rc = gemm_init()) return rc;
    gemm_set_cta_group(cta_group);
    ++g_switch_epoch;
    return 0;
}

int pafuse_set_fuse_layernorm(pafuse_ctx* ctx, int32_t enable) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    ctx->fuse_ln = enable != 0;
    drop_graphs(ctx);
    return 0;
}

int pafuse_set_fuse_mlp(pafuse_ctx* ctx, int32_t enable) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    ctx->fuse_mlp = enable != 0;
    drop_graphs(ctx);
    return 0;
}

int pafuse_set_part_streams(pafuse_ctx* ctx, int32_t enable, const int32_t* shares) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    ctx->part_streams = enable != 0;
    drop_graphs(ctx);
    if (shares) {
        for (int pi = 0; pi < ctx->num_parts; ++pi) {
            if (shares[pi] < 2) {
                set_last_error("pafuse_set_part_streams: a share needs at least 2 SMs (part %d: %d)", pi, shares[pi]);
                return PAFUSE_E_ARG;
            }
            ctx->sm_share[pi] = shares[pi] / 2 * 2;
        }
        ctx->shares_fixed = true;
    }
    return 0;
}

int pafuse_set_gemm_weight_stationary(int32_t enable) {
    if (int rc = gemm_init()) return rc;
    gemm_set_weight_stationary(enable != 0);
    ++g_switch_epoch;
    return 0;
}

int pafuse_profile_enable(pafuse_ctx* ctx, int32_t enable) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    Profiler& pr = ctx->prof;
    for (ProfRec& r : pr.recs) {
        pr.pool.push_back(r.a);
        pr.pool.push_back(r.b);
    }
    pr.recs.clear();
    pr.on = enable != 0;
    return 0;
}

int pafuse_profile_read(pafuse_ctx* ctx, double* ms, double* work, int64_t* launches, int32_t ncat) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!ms || !work || !launches || ncat < CAT_COUNT) {
        set_last_error("pafuse_profile_read: need room for %d categories", (int)CAT_COUNT);
        return PAFUSE_E_ARG;
    }
    for (int i = 0; i < ncat; ++i) {
        ms[i] = 0.0;
        work[i] = 0.0;
        launches[i] = 0;
    }
    Profiler& pr = ctx->prof;
    for (ProfRec& r : pr.recs) {
        PAFUSE_CUDA_OK(cudaEventSynchronize(r.b));
        float t = 0.f;
        PAFUSE_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cat] += (double)t;
        work[r.cat] += r.work;
        launches[r.cat] += 1;
    }
    // PAFUSE_PROF_DETAIL=1: the same launches grouped by (category, algorithmic work, bytes) -- one line per kernel kind
    // and part, timed INSIDE the step (the ncu launch lists time every kernel alone, at another clock)
    if (getenv("PAFUSE_PROF_DETAIL")) {
        static const char* names[CAT_COUNT] = {"gemm", "attention", "layernorm", "embed_head", "ddim", "post"};
        std::map<std::tuple<int, double, double>, std::pair<double, long long>> groups;
        for (ProfRec& r : pr.recs) {
            float t = 0.f;
            cudaEventElapsedTime(&t, r.a, r.b);
            auto& g = groups[std::make_tuple(r.cat, r.work, r.bytes)];
            g.first += (double)t;
            g.second += 1;
        }
        for (auto& kv : groups)
            fprintf(stderr, "prof_detail %-10s work %.4e bytes %.4e launches %6lld avg_us %9.1f total_ms %9.2f\n", names[std::get<0>(kv.first)],
                    std::get<1>(kv.first), std::get<2>(kv.first), kv.second.second, kv.second.first / (double)kv.second.second * 1e3,
                    kv.second.first);
    }
    return 0;
}

int pafuse_profile_read_bytes(pafuse_ctx* ctx, double* bytes, int32_t ncat) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!bytes || ncat < CAT_COUNT) {
        set_last_error("pafuse_profile_read_bytes: need room for %d categories", (int)CAT_COUNT);
        return PAFUSE_E_ARG;
    }
    for (int i = 0; i < ncat; ++i) bytes[i] = 0.0;
    for (ProfRec& r : ctx->prof.recs) bytes[r.cat] += r.bytes;
    return 0;
}

// ---- unit-level entry points -------------------------------------------------------------

__global__ void split_rows_kernel(const float* x, op_t* hi, op_t* lo, size_t n, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_op(x[i] * scale, hi[i], lo[i]);
}
__global__ void join_rows_kernel(const op_t* hi, const op_t* lo, float* y, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = join_op(hi[i], lo[i]);
}

int pafuse_linear(pafuse_ctx* ctx, const float* x, const float* w, const float* b, float* y, int64_t M, int32_t N,
                  int32_t K, int32_t epilogue, int32_t use_simt, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!x || !w || !b || !y || M < 1 || N < 1 || K < 1 || epilogue < 0 || epilogue > 2) {
        set_last_error("pafuse_linear: bad argument");
        return PAFUSE_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    op_t *xh = nullptr, *xl = nullptr, *wh = nullptr, *wl = nullptr, *yh = nullptr, *yl = nullptr;
    size_t nx = (size_t)M * K, nw = (size_t)N * K, ny = (size_t)M * N;
    int rc = 0;
    if (dev_alloc(&xh, nx) || dev_alloc(&xl, nx) || dev_alloc(&wh, nw) || dev_alloc(&wl, nw)) rc = PAFUSE_E_CUDA;
    if (!rc && epilogue == EPI_GELU_SPLIT && (dev_alloc(&yh, ny) || dev_alloc(&yl, ny))) rc = PAFUSE_E_CUDA;
    if (!rc) {
        // weights first: the GEMM may fetch its W slice ahead of its grid-dependency wait, which only covers the
        // kernel launched immediately before it
        split_rows_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(w, wh, wl, nw, WEIGHT_SCALE);
        split_rows_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(x, xh, xl, nx, 1.0f);
        GemmArgs g;
        g.a_hi = xh; g.a_lo = xl; g.w_hi = wh; g.w_lo = wl; g.bias = b; g.out_f32 = y; g.out_hi = yh; g.out_lo = yl;
        g.M = M; g.N = N; g.K = K; g.epilogue = epilogue;
        rc = use_simt ? launch_gemm_simt(g, st) : launch_gemm_tcgen05(g, st);
        if (!rc && epilogue == EPI_GELU_SPLIT) join_rows_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, st>>>(yh, yl, y, ny);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(xh); cudaFree(xl); cudaFree(wh); cudaFree(wl); cudaFree(yh); cudaFree(yl);
    if (!rc && e != cudaSuccess) {
        set_last_error("pafuse_linear: %s", cudaGetErrorString(e));
        rc = PAFUSE_E_CUDA;
    }
    return rc;
}

int pafuse_attention(pafuse_ctx* ctx, const float* qkv, float* out, int32_t S, int32_t J, int32_t C, int32_t temporal,
                     void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!qkv || !out || S < 1) {
        set_last_error("pafuse_attention: bad argument");
        return PAFUSE_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long M = (long long)S * ctx->cfg.frames * J;
    size_t n = (size_t)M * C;
    op_t *oh = nullptr, *ol = nullptr, *ph = nullptr, *pl = nullptr;
    int rc = 0;
    if (dev_alloc(&oh, n) || dev_alloc(&ol, n)) rc = PAFUSE_E_CUDA;
    if (!rc) {
        if (ctx->debug_simt_attn) {
            AttnParams a;
            a.qkv = qkv; a.out_hi = oh; a.out_lo = ol; a.S = S; a.F = ctx->cfg.frames; a.J = J; a.C = C;
            a.temporal = temporal; a.scale = 0.f;
            rc = launch_attention(a, st);
        } else {
            AttnPlanes planes;
            planes.hds = attn_head_store(C / 8);
            planes.rows_cap = M;
            size_t halves = (size_t)24 * M * planes.hds;
            if (dev_alloc(&ph, halves) || dev_alloc(&pl, halves)) rc = PAFUSE_E_CUDA;
            planes.hi = ph; planes.lo = pl;
            if (!rc) rc = launch_qkv_to_planes(qkv, planes, M, C, st);
            if (!rc) rc = launch_attention_tc(planes, oh, ol, S, ctx->cfg.frames, J, C, temporal, st);
        }
        if (!rc) join_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oh, ol, out, n);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(oh); cudaFree(ol); cudaFree(ph); cudaFree(pl);
    if (!rc && e != cudaSuccess) {
        set_last_error("pafuse_attention: %s", cudaGetErrorString(e));
        rc = PAFUSE_E_CUDA;
    }
    return rc;
}

int pafuse_qkv_attention(pafuse_ctx* ctx, const float* x, const float* w, const float* b, float* out, int32_t S,
                         int32_t J, int32_t C, int32_t temporal, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!x || !w || !b || !out || S < 1 || J < 1 || C < 8 || C % 8 != 0) {
        set_last_error("pafuse_qkv_attention: bad argument");
        return PAFUSE_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long M = (long long)S * ctx->cfg.frames * J;
    const int hd = C / 8, hds = attn_head_store(hd);
    const size_t n = (size_t)M * C, nw = (size_t)24 * hds * C, halves = (size_t)24 * M * hds;
    op_t *xh = nullptr, *xl = nullptr, *wh = nullptr, *wl = nullptr, *ph = nullptr, *pl = nullptr, *oh = nullptr, *ol = nullptr;
    float *wp = nullptr, *bp = nullptr;
    int rc = 0;
    if (dev_alloc(&xh, n) || dev_alloc(&xl, n) || dev_alloc(&wh, nw) || dev_alloc(&wl, nw) || dev_alloc(&ph, halves) ||
        dev_alloc(&pl, halves) || dev_alloc(&oh, n) || dev_alloc(&ol, n) || dev_alloc(&wp, nw) ||
        dev_alloc(&bp, (size_t)24 * hds))
        rc = PAFUSE_E_CUDA;
    if (!rc) rc = launch_pack_qkv(w, b, wp, bp, C, hd, hds, st);
    if (!rc) {
        split_rows_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(wp, wh, wl, nw, WEIGHT_SCALE);   // weights first, see pafuse_linear
        split_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, xh, xl, n, 1.0f);
        GemmArgs g;
        g.a_hi = xh; g.a_lo = xl; g.w_hi = wh; g.w_lo = wl; g.bias = bp; g.out_f32 = nullptr; g.out_hi = g.out_lo = nullptr;
        g.M = M; g.N = 24 * hds; g.K = C; g.epilogue = EPI_PLANES;
        g.planes.hi = ph; g.planes.lo = pl; g.planes.rows_cap = M; g.planes.hds = hds;
        rc = ctx->debug_simt ? launch_gemm_simt(g, st) : launch_gemm_tcgen05(g, st);
        if (!rc) rc = launch_attention_tc(g.planes, oh, ol, S, ctx->cfg.frames, J, C, temporal, st);
        if (!rc) join_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oh, ol, out, n);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(xh); cudaFree(xl); cudaFree(wh); cudaFree(wl); cudaFree(ph); cudaFree(pl); cudaFree(oh); cudaFree(ol);
    cudaFree(wp); cudaFree(bp);
    if (!rc && e != cudaSuccess) {
        set_last_error("pafuse_qkv_attention: %s", cudaGetErrorString(e));
        rc = PAFUSE_E_CUDA;
    }
    return rc;
}

int pafuse_mlp_block(pafuse_ctx* ctx, const float* a, const float* w1, const float* b1, const float* w2, const float* b2,
                     float* x, const float* g0, const float* bb0, const float* g1, const float* bb1, float* a_out, int64_t M,
                     int32_t C, int32_t fused, void* stream) {
    if (!check_ctx(ctx)) return PAFUSE_E_ARG;
    if (!a || !w1 || !b1 || !w2 || !b2 || !x || !g1 || !bb1 || !a_out || (g0 && !bb0) || M < 1 || !mlp_can_fuse(C)) {
        set_last_error("pafuse_mlp_block: bad argument (C %% 32 == 0, 64 <= C <= 256)");
        return PAFUSE_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)M * C, nw = (size_t)2 * C * C;
    op_t *ah = nullptr, *al = nullptr, *w1h = nullptr, *w1l = nullptr, *w2h = nullptr, *w2l = nullptr, *hh = nullptr,
         *hl = nullptr, *oh = nullptr, *ol = nullptr;
    int rc = 0;
    if (dev_alloc(&ah, n) || dev_alloc(&al, n) || dev_alloc(&w1h, nw) || dev_alloc(&w1l, nw) || dev_alloc(&w2h, nw) ||
        dev_alloc(&w2l, nw) || dev_alloc(&oh, n) || dev_alloc(&ol, n) || (!fused && (dev_alloc(&hh, 2 * n) || dev_alloc(&hl, 2 * n))))
        rc = PAFUSE_E_CUDA;
    if (!rc) {
        split_rows_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(w1, w1h, w1l, nw, WEIGHT_SCALE);   // weights first, see pafuse_linear
        split_rows_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(w2, w2h, w2l, nw, WEIGHT_SCALE);
        split_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, ah, al, n, 1.0f);
        GemmLnFuse ln;
        ln.x = x; ln.g0 = g0; ln.b0 = bb0; ln.g1 = g1; ln.b1 = bb1; ln.eps0 = 1e-6f; ln.eps1 = 1e-6f;
        if (fused) {
            MlpArgs m;
            m.a_hi = ah; m.a_lo = al; m.w1_hi = w1h; m.w1_lo = w1l; m.b1 = b1; m.w2_hi = w2h; m.w2_lo = w2l; m.b2 = b2;
            m.x = x; m.out_hi = oh; m.out_lo = ol; m.M = M; m.C = C; m.ln = ln;
            rc = launch_mlp_fused(m, st);
        } else {
            GemmArgs g;
            g.a_hi = ah; g.a_lo = al; g.w_hi = w1h; g.w_lo = w1l; g.bias = b1; g.out_f32 = nullptr; g.out_hi = hh; g.out_lo = hl;
            g.M = M; g.N = 2 * C; g.K = C; g.epilogue = EPI_GELU_SPLIT;
            rc = launch_gemm_tcgen05(g, st);
            g = GemmArgs();
            g.a_hi = hh; g.a_lo = hl; g.w_hi = w2h; g.w_lo = w2l; g.bias = b2; g.out_f32 = x; g.out_hi = oh; g.out_lo = ol;
            g.M = M; g.N = C; g.K = 2 * C; g.epilogue = EPI_RESID_LN; g.ln = ln;
            if (!rc) rc = launch_gemm_tcgen05(g, st);
        }
        if (!rc) join_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oh, ol, a_out, n);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(ah); cudaFree(al); cudaFree(w1h); cudaFree(w1l); cudaFree(w2h); cudaFree(w2l); cudaFree(hh); cudaFree(hl);
    cudaFree(oh); cudaFree(ol);
    if (!rc && e != cudaSuccess) {
        set_last_error("pafuse_mlp_block: %s", cudaGetErrorString(e));
        rc = PAFUSE_E_CUDA;
    }
    return rc;
}

}  // extern "C"
