"""GPU (-m gpu): the sm_100a path, called through the C ABI, against the CPU oracle / golden fixtures.

Tolerances (BASELINE.json north_star): floating point within 1e-3 relative per coordinate and
<= 0.1 mm MPJPE delta; coordinates are metres, and "relative" is evaluated as
|d| <= 1e-3*|ref| + 2e-5 (the absolute term, 0.02 mm, covers coordinates that are ~0, e.g. near the
root).  Integer / index work (part tables, flip permutation, re-assembly, J-Agg selection) is bit-exact.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL, MPJPE_MM = 1e-3, 2e-5, 0.1


def _ctx():
    from pafuse_b200 import _native
    return _native.NativeContext(27, 134, 1, 8, [32], [[0]], list(range(134)), 1.0, 1, torch.device("cuda", 0))


def _model(c, simt=False):
    import pafuse_b200
    from pafuse_b200.h3wb import H3WBSkeleton
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(c["args"], sk.joints_left, sk.joints_right, sk, is_train=False, num_proposals=c["H"],
                         sampling_timesteps=c["K"])
    m.load_state_dict(c["sd"], strict=False)
    m = m.cuda().eval()
    noises = c["noises"]
    m.noise_source = lambda k, shape, device: noises[k].to(device)
    if simt:
        m.native_context().set_debug_simt_gemm(True)
    return m


def _check_pose(out, ref):
    d = out.double().cpu() - ref.double()
    mpjpe = d.norm(dim=-1).mean().item() * 1e3
    assert mpjpe <= MPJPE_MM, f"MPJPE delta {mpjpe} mm"
    bad = d.abs() > RTOL * ref.double().abs() + ATOL
    assert not bad.any(), f"{int(bad.sum())} coordinates out of tolerance, max abs {d.abs().max().item():.3e}"
    return mpjpe


# ------------------------------------------------------------------ unit level
@pytest.mark.parametrize("N,K", [(1152, 384), (384, 384), (768, 384), (384, 768), (672, 224), (224, 224), (448, 224),
                                 (224, 448), (768, 256), (256, 256), (512, 256), (256, 512)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_linear_shapes_of_the_three_parts(N, K, epi):
    torch.manual_seed(N * 7 + K + epi)
    c = _ctx()
    M = 148 * 128 + 77                                                # more than one persistent wave + a ragged tile
    x = torch.randn(M, K, device="cuda")
    w = (torch.rand(N, K, device="cuda") * 2 - 1) / K ** 0.5
    b = torch.randn(N, device="cuda") * 0.1
    ref = x.double() @ w.double().t() + b.double()
    y = None
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        y0 = torch.randn(M, N, device="cuda")
        ref, y = ref + y0.double(), y0.clone()
    out = c.linear(x, w, b, epilogue=epi, y=y)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item()
    assert err < 5e-5, err
    fp32 = torch.nn.functional.linear(x, w, b)                        # torch fp32 (no TF32) is itself ~1e-6 from fp64
    assert err < 40 * max((fp32.double() - (x.double() @ w.double().t() + b.double())).abs().max().item(), 1e-6)


@pytest.mark.parametrize("M", [1, 127, 128, 129])
def test_linear_ragged_rows(M):
    torch.manual_seed(M)
    c = _ctx()
    x, w, b = torch.randn(M, 256, device="cuda"), torch.randn(512, 256, device="cuda") / 16, torch.randn(512, device="cuda")
    out = c.linear(x, w, b)
    ref = x.double() @ w.double().t() + b.double()
    assert (out.double() - ref).abs().max().item() < 5e-5


@pytest.mark.parametrize("C", [64, 128, 224, 256])
@pytest.mark.parametrize("M", [1, 255, 256, 257, 74 * 256 + 513])
@pytest.mark.parametrize("chained", [False, True])
def test_mlp_block_one_kernel_against_two_launches_and_fp64(C, M, chained):
    """pafuse_mlp_block: x + fc2(GELU(fc1(a))) and the LayerNorms that follow, as ONE kernel (hidden activations in
    tensor memory) against the fc1 / fc2 GEMM launches (same MMAs in the same order) and against fp64 torch; ragged
    row counts, a tile per CTA pair and several, hidden widths with and without a 64-column last chunk."""
    import torch.nn.functional as Fn
    torch.manual_seed(C + M)
    c = _ctx()
    dev = "cuda"
    a = torch.randn(M, C, device=dev)
    x = torch.randn(M, C, device=dev)
    w1 = (torch.rand(2 * C, C, device=dev) * 2 - 1) / C ** 0.5
    w2 = (torch.rand(C, 2 * C, device=dev) * 2 - 1) / (2 * C) ** 0.5
    b1, b2 = torch.randn(2 * C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
    g1, bb1 = 1 + 0.1 * torch.randn(C, device=dev), 0.1 * torch.randn(C, device=dev)
    g0, bb0 = (1 + 0.1 * torch.randn(C, device=dev), 0.1 * torch.randn(C, device=dev)) if chained else (None, None)
    xf, af = c.mlp_block(a, w1, b1, w2, b2, x, g1, bb1, g0, bb0, fused=True)
    xs, as_ = c.mlp_block(a, w1, b1, w2, b2, x, g1, bb1, g0, bb0, fused=False)
    assert (xf - xs).abs().max().item() < 1e-5 and (af - as_).abs().max().item() < 1e-5
    d = lambda t: t.double()
    v = d(x) + Fn.linear(Fn.gelu(Fn.linear(d(a), d(w1), d(b1))), d(w2), d(b2))
    if chained:
        v = Fn.layer_norm(v, (C,), d(g0), d(bb0), 1e-6)
    ref_a = Fn.layer_norm(v, (C,), d(g1), d(bb1), 1e-6)
    assert (d(xf) - v).abs().max().item() < 3e-5
    assert (d(af) - ref_a).abs().max().item() < 3e-5


@pytest.mark.parametrize("N,K", [(1152, 384), (384, 384), (768, 384), (768, 224), (224, 224), (448, 224), (768, 256),
                                 (256, 256), (512, 256), (32, 64)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_weight_stationary_tiles_equal_streamed_tiles(N, K, epi):
    """Both GEMM schedules issue the same MMA sequence per output tile: results must be bit-identical."""
    torch.manual_seed(N + K + epi)
    c = _ctx()
    M = 74 * 256 * 2 + 300                                            # several m tiles per CTA pair + a ragged one
    x = torch.randn(M, K, device="cuda")
    w = (torch.rand(N, K, device="cuda") * 2 - 1) / K ** 0.5
    b = torch.randn(N, device="cuda") * 0.1
    y0 = torch.randn(M, N, device="cuda")
    try:
        c.set_gemm_weight_stationary(False)
        streamed = c.linear(x, w, b, epilogue=epi, y=y0.clone())
        c.set_gemm_weight_stationary(True)
        resident = c.linear(x, w, b, epilogue=epi, y=y0.clone())
    finally:
        c.set_gemm_weight_stationary(True)
    torch.cuda.synchronize()
    assert torch.equal(streamed, resident)      # epi 2: x += y is one L2-side add per element in both schedules
    ref = x.double() @ w.double().t() + b.double()
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = ref + y0.double()
    assert (resident.double() - ref).abs().max().item() < 5e-5


@pytest.mark.parametrize("J,C", [(24, 384), (68, 224), (42, 256), (17, 256), (30, 384), (5, 224)])
@pytest.mark.parametrize("temporal", [False, True])
def test_attention(J, C, temporal):
    torch.manual_seed(J + C)
    c = _ctx()
    S, F, hd = 3, 27, C // 8
    qkv = torch.randn(S * F * J, 3 * C, device="cuda")
    out = c.attention(qkv, S, J, C, temporal)
    t = qkv.double().reshape(S, F, J, 3, 8, hd)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    perm = (0, 2, 3, 1, 4) if temporal else (0, 1, 3, 2, 4)
    q, k, v = (z.permute(*perm) for z in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v
    a = a.permute(0, 3, 1, 2, 4) if temporal else a.permute(0, 1, 3, 2, 4)
    assert (out.double() - a.reshape(S * F * J, C)).abs().max().item() < 5e-5


def test_single_denoiser_module_against_oracle():
    """MixSTE2 sub-boundary (mixste.py:278) with randomised positional embeddings."""
    import pafuse_b200
    from oracle import pafuse_oracle as orc
    from pafuse_b200 import synthetic
    sd = synthetic.synthetic_state_dict(seed=4, depth=2)
    w = orc.part_weights(sd, "face")
    net = pafuse_b200.MixSTE2(num_frame=27, num_joints=68, in_chans=5, embed_dim_ratio=224, depth=2, num_heads=8,
                              mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_path_rate=0, is_train=False)
    net.load_state_dict(w)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(2)
    x2d, x3d = torch.rand(2, 27, 68, 2, generator=g) * 2 - 1, torch.randn(2, 3, 27, 68, 3, generator=g)
    t = torch.full((2,), 399, dtype=torch.long)
    got = net(x2d.cuda(), x3d.cuda(), t.cuda())
    _check_pose(got, orc.mixste_forward(w, x2d, x3d, t, depth=2))


# ------------------------------------------------------------------ sampler against the golden fixtures
@pytest.mark.parametrize("name", ["tiny_B1_H3_K2_d2", "noflip_B2_H1_K2_d2", "cfg1_B2_H1_K1", "small_B2_H2_K3"])
def test_sampler_matches_reference_golden(name):
    from pafuse_testlib import build_case
    c = build_case(name)
    m = _model(c)
    out = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda() if c["flip"] else None)
    assert out.shape == c["golden"]["out"].shape and out.dtype == torch.float32
    _check_pose(out, c["golden"]["out"])


def _compare_bench_shape(out, c):
    """Sampler output against a bench-shape fixture (tests/golden/make_golden_bench_shapes.py): every step in full
    ("out"), or the last step in full + three frames of every earlier step ("out_last" / "out_frames")."""
    g = c["golden"]
    if "out" in g:
        return _check_pose(out, g["out"])
    keep = [int(v) for v in c["g"]["keep_frames"]]
    _check_pose(out[:, :-1][:, :, :, keep], g["out_frames"])
    return _check_pose(out[:, -1], g["out_last"])


def test_sampler_at_the_benchmarked_shape_H5_K5():
    """BASELINE.json configs[1] (the bench.py headline): num_proposals=5, sampling_timesteps=5, depth 8, flip-TTA,
    against the reference's own output for the same clip and the same injected noise (diffusionpose.py:272-316)."""
    from pafuse_testlib import build_case
    c = build_case("cfg2_B1_H5_K5")
    m = _model(c)
    out = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert out.shape == (1, 5, 5, 27, 134, 3)
    _compare_bench_shape(out, c)


def test_sampler_at_the_hypothesis_sharded_shape_H20_K10():
    """BASELINE.json configs[2]: num_proposals=20, sampling_timesteps=10 (ten chained DDIM updates)."""
    from pafuse_testlib import build_case
    c = build_case("cfg3_B1_H20_K10")
    m = _model(c)
    out = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert out.shape == (1, 10, 20, 27, 134, 3)
    _compare_bench_shape(out, c)


def test_bench_shape_clip_inside_a_64_clip_batch():
    """The clip of the H=5, K=5 fixture lifted as clip 0 of the 64-clip bench batch (the other 63 clips get other
    inputs and noise): rows are independent, so its result must still match the reference's single-clip output."""
    from pafuse_b200 import synthetic
    from pafuse_testlib import build_case
    c = build_case("cfg2_B1_H5_K5")
    B = 64
    x2d, x2df = synthetic.synthetic_inputs(B, seed=9)
    x2d[0], x2df[0] = c["x2d"][0], c["x2df"][0]
    noises = synthetic.synthetic_noise(B, c["H"], c["K"], seed=9)
    for k in range(c["K"]):
        noises[k][0] = c["noises"][k][0]
    m = _model(c)
    m.noise_source = lambda k, shape, device: noises[k].to(device)
    out = m(x2d.cuda(), None, input_2d_flip=x2df.cuda())
    _check_pose(out[:1], c["golden"]["out"])


def test_tensor_core_gemm_equals_cuda_core_gemm_on_the_whole_model():
    from pafuse_testlib import build_case
    c = build_case("tiny_B1_H3_K2_d2")
    a = _model(c)(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    b = _model(c, simt=True)(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert (a - b).abs().max().item() < 3e-5


def test_layernorm_fused_in_gemm_epilogue_equals_separate_launches():
    """proj / fc2 with the following LayerNorms in their epilogue (face, hands) against the ln_chain launches."""
    from pafuse_testlib import build_case
    c = build_case("small_B2_H2_K3")
    fused = _model(c)
    plain = _model(c)
    plain.native_context().set_fuse_layernorm(False)
    a = fused(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    b = plain(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert (a - b).abs().max().item() < 2e-5
    _check_pose(a, c["golden"]["out"])
    _check_pose(b, c["golden"]["out"])


def test_mlp_fused_in_one_kernel_equals_fc1_fc2_launches():
    """fc1 + GELU + fc2 + residual + LayerNorms as one kernel (face, hands; hidden activations in tensor memory)
    against the two GEMM launches: the same MMAs in the same order on the same operands."""
    from pafuse_testlib import build_case
    c = build_case("small_B2_H2_K3")
    fused = _model(c)
    fused.native_context().set_fuse_mlp(True)
    plain = _model(c)
    plain.native_context().set_fuse_mlp(False)
    a = fused(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    b = plain(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert (a - b).abs().max().item() < 2e-5
    _check_pose(a, c["golden"]["out"])


@pytest.mark.parametrize("shares", [None, (20, 96, 32), (2, 2, 2)])
def test_parts_side_by_side_equal_parts_in_turn(shares):
    """The part denoisers on their own streams and SM shares (opt-in) against one after the other on the
    caller's stream: every tile is computed by the same code in the same order, so the result is bit-identical,
    whatever the shares."""
    from pafuse_testlib import build_case
    c = build_case("small_B2_H2_K3")
    side = _model(c)
    side.native_context().set_part_streams(True, shares)
    turn = _model(c)
    turn.native_context().set_part_streams(False)
    a = side(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    b = turn(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert torch.equal(a, b)
    _check_pose(a, c["golden"]["out"])


@pytest.mark.parametrize("max_seqs", [1, 3])
def test_workspace_chunking_is_invisible(max_seqs):
    """max_seqs smaller than the sequence count must not change the result.  Rows are independent, but the
    tcgen05 attention packs several short sequences into one 128-row tile, so the position of a sequence in
    its tile (hence the order in which the tensor core adds the zero contributions of the other groups) moves
    with the chunk boundaries: equal to fp32 summation-order noise, not bit-equal."""
    from pafuse_testlib import build_case
    c = build_case("small_B2_H2_K3")
    full = _model(c)(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    m = _model(c)
    m.max_seqs = max_seqs
    m._native_dirty = True
    chunked = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert (full - chunked).abs().max().item() < 5e-6
    rerun = _model(c)(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert torch.equal(full, rerun)                                   # same launch geometry -> bit-identical


@pytest.mark.parametrize("J,C", [(24, 384), (68, 224), (42, 256)])
@pytest.mark.parametrize("temporal", [False, True])
def test_qkv_gemm_head_plane_epilogue_feeds_attention(J, C, temporal):
    """qkv GEMM (head-plane epilogue) + tcgen05 attention == softmax(q k^T / sqrt(hd)) v of mixste.py:63-79."""
    torch.manual_seed(J * 3 + C + int(temporal))
    c = _ctx()
    S, F, hd = 7, 27, C // 8
    M = S * F * J
    x = torch.randn(M, C, device="cuda")
    w = (torch.rand(3 * C, C, device="cuda") * 2 - 1) / C ** 0.5 * 2
    b = torch.randn(3 * C, device="cuda") * 0.1
    out = c.qkv_attention(x, w, b, S, J, C, temporal)
    qkv = x.double() @ w.double().t() + b.double()
    t = qkv.reshape(S, F, J, 3, 8, hd)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    perm = (0, 2, 3, 1, 4) if temporal else (0, 1, 3, 2, 4)
    q, k, v = (z.permute(*perm) for z in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v
    a = a.permute(0, 3, 1, 2, 4) if temporal else a.permute(0, 1, 3, 2, 4)
    assert (out.double() - a.reshape(M, C)).abs().max().item() < 5e-5


def test_default_noise_path_is_deterministic_and_finite():
    from pafuse_testlib import build_case
    c = build_case("tiny_B1_H3_K2_d2")
    m = _model(c)
    m.noise_source = None
    torch.manual_seed(1)
    a = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    torch.manual_seed(1)
    b = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda())
    assert torch.isfinite(a).all() and torch.equal(a, b)
    assert a.abs().max().item() <= 1.1 + 1e-6                         # x0 is clamped to +-1.1*scale


def test_empty_batch():
    from pafuse_testlib import build_case
    c = build_case("tiny_B1_H3_K2_d2")
    m = _model(c)
    m.noise_source = None
    out = m(c["x2d"][:0].cuda(), None, input_2d_flip=c["x2df"][:0].cuda())
    assert out.shape == (0, c["K"], c["H"], 27, 134, 3)


# ------------------------------------------------------------------ post-processing (bit-exact parts)
@pytest.mark.parametrize("name", ["tiny_B1_H3_K2_d2", "small_B2_H2_K3"])
def test_reassembly_bit_exact_including_input_mutation(name):
    import pafuse_b200
    from pafuse_b200.h3wb import H3WBSkeleton
    from pafuse_testlib import build_case
    c = build_case(name)
    x = c["golden"]["out"].cuda().clone()
    ds = H3WBSkeleton()
    wb = pafuse_b200.wb_pose_from_parts(x, ds)
    assert torch.equal(wb.cpu(), c["golden"]["wb"])
    assert torch.equal(x.cpu(), c["golden"]["wb_input_after"])        # rows 0/1/10/11 negated like the reference
    assert not torch.signbit(wb[..., 0, :]).any()
    assert ds.parts_connection_indices["body"] == 0                   # utils.py:116 side effect


@pytest.mark.parametrize("name", ["tiny_B1_H3_K2_d2", "small_B2_H2_K3", "cfg1_B2_H1_K1"])
def test_projection_and_aggregation_against_golden(name):
    import pafuse_b200
    from pafuse_testlib import build_case
    c = build_case(name)
    g = c["golden"]
    wb = g["wb"].cuda()
    B, K, H, F, J, _ = wb.shape
    absd = (g["wb"] + c["traj"][:, None, None]).reshape(B * K * H * F, J, 3)
    rep = pafuse_b200.project_to_2d(absd.cuda(), c["cam"].repeat(B * K * H * F, 1).cuda())
    assert torch.equal(rep.cpu().reshape(B, K, H, F, J, 2), g["reproj"])
    jagg, pagg, sel, rep2 = pafuse_b200.aggregate_hypotheses(wb, c["traj"].cuda(), c["cam"].cuda(), c["x2d"].cuda(),
                                                             return_select=True, return_reproj=True)
    assert torch.equal(rep2.cpu(), g["reproj"])
    assert torch.equal(sel.cpu().long(), g["select"])
    assert torch.equal(jagg.cpu(), g["jagg"])
    assert torch.allclose(pagg.cpu(), g["pagg"], rtol=0, atol=2.4e-7)    # mean: same values, summation order may differ by 1 ulp


def test_aggregation_properties_at_full_size():
    """BASELINE config 2 size (B=64,K=5,H=5): size-independent properties instead of a CPU re-computation."""
    import pafuse_b200
    from pafuse_b200 import synthetic
    B, K, H = 64, 5, 5
    g = torch.Generator().manual_seed(0)
    wb = (torch.randn(B, K, H, 27, 134, 3, generator=g) * 0.3).cuda()
    x2d = synthetic.synthetic_inputs(B, seed=2)[0].cuda()
    traj, cam = synthetic.synthetic_trajectory(B, seed=2).cuda(), synthetic.h36m_cam0_intrinsics().cuda()
    jagg, pagg, sel = pafuse_b200.aggregate_hypotheses(wb, traj, cam, x2d, return_select=True)
    assert sel.min().item() >= 0 and sel.max().item() < H
    picked = torch.gather(wb, 2, sel.long()[:, :, None, :, :, None].expand(B, K, 1, 27, 134, 3)).squeeze(2)
    assert torch.equal(picked, jagg)                                  # J-Agg returns an actual hypothesis, bit for bit
    assert torch.allclose(pagg, wb.mean(dim=2), rtol=0, atol=3e-7)
    perm = torch.tensor([3, 1, 4, 0, 2], device="cuda")               # permuting hypotheses permutes the selection
    j2, p2, s2 = pafuse_b200.aggregate_hypotheses(wb[:, :, perm].contiguous(), traj, cam, x2d, return_select=True)
    assert torch.equal(j2, jagg)
    same = wb[:, :, :1].expand(B, K, H, 27, 134, 3).contiguous()      # identical hypotheses: first index wins, mean == value
    j3, p3, s3 = pafuse_b200.aggregate_hypotheses(same, traj, cam, x2d, return_select=True)
    assert int(s3.abs().sum()) == 0 and torch.equal(j3, same[:, :, 0])


# ------------------------------------------------------------------ end to end and sharding
def test_lift_end_to_end_against_oracle():
    from pafuse_b200 import distributed as pd
    from pafuse_b200.h3wb import H3WBSkeleton
    from pafuse_testlib import build_case
    c = build_case("small_B2_H2_K3")
    m = _model(c)
    eng = pd.CudaEngine(m, H3WBSkeleton())
    noises = c["noises"]
    res = pd.lift(eng, c["x2d"].cuda(), c["x2df"].cuda(), c["traj"].cuda(), c["cam"].cuda(), c["H"],
                  noise_source=lambda k, shape, device: noises[k].to(device), keep_hypotheses=True)
    _check_pose(res.pred, c["golden"]["wb"])
    _check_pose(res.pagg, c["golden"]["pagg"])
    agree = (res.select.cpu().long() == c["golden"]["select"]).float().mean().item()
    assert agree > 0.995                                              # picks may differ only where two errors tie to ~1e-6


def test_sharded_equals_single_device():
    """lift_sharded with (rank, world) given explicitly: the union of the shards equals the 1-GPU result bit for bit."""
    from pafuse_b200 import distributed as pd
    from pafuse_b200.h3wb import H3WBSkeleton
    from pafuse_testlib import build_case
    c = build_case("tiny_B1_H3_K2_d2")
    from pafuse_b200 import synthetic
    B, H = 3, 3
    x2d, x2df = synthetic.synthetic_inputs(B, seed=6)
    traj, cam = synthetic.synthetic_trajectory(B, seed=6).cuda(), synthetic.h36m_cam0_intrinsics().cuda()
    m = _model(c)
    eng = pd.CudaEngine(m, H3WBSkeleton())
    one = pd.lift_sharded(eng, x2d.cuda(), x2df.cuda(), traj, cam, H, mode="clips", seed=3, rank=0, world=1)
    parts = [pd.lift_sharded(eng, x2d.cuda(), x2df.cuda(), traj, cam, H, mode="clips", seed=3, rank=r, world=2,
                             gather=False) for r in range(2)]
    assert torch.equal(torch.cat([p.jagg for p in parts]), one.jagg)
    assert torch.equal(torch.cat([p.pagg for p in parts]), one.pagg)
    assert torch.equal(torch.cat([p.select for p in parts]), one.select)
