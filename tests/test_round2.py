"""Round-2 additions: counter-based sampler noise, CUDA-graph replay of small passes, weight tracking, per-device kernel
attributes, nn.DataParallel tolerance, operand range check, the in-the-wild .npy output.  CPU tests first, then -m gpu."""
import os

import numpy as np
import pytest
import torch

from oracle import philox_ref


# ------------------------------------------------------------------ CPU: the generator's definition
def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    w = philox_ref.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in w] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    w = philox_ref.philox4x32_10([f], [f], [f], [f], f, f)
    assert [int(x[0]) for x in w] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    w = philox_ref.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in w] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normals_are_standard_and_position_addressed():
    x = philox_ref.randn(seed=3, draw=0, base=0, n=400000)
    assert abs(float(x.mean())) < 5e-3 and abs(float(x.std()) - 1) < 5e-3 and np.isfinite(x).all()
    assert abs(float((x ** 4).mean()) - 3.0) < 0.1                    # kurtosis of a Gaussian
    part = philox_ref.randn(seed=3, draw=0, base=12345, n=1000)       # any slice, even at an odd offset, is the same values
    assert np.array_equal(part, x[12345:13345])
    other = philox_ref.randn(seed=3, draw=1, base=0, n=1000)
    assert abs(float(np.corrcoef(other, x[:1000])[0, 1])) < 0.15      # another draw is another stream


def test_weight_fingerprint_tracks_in_place_edits_and_child_loads():
    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    from pafuse_b200.mixste import weights_fingerprint
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(synthetic.default_args(depth=1), sk.joints_left, sk.joints_right, sk, is_train=False)
    fp0 = m._fingerprint()
    assert m._fingerprint() == fp0
    with torch.no_grad():
        m.pose_estimator["face"].head[1].bias.add_(1.0)               # in-place edit of a child's parameter
    fp1 = m._fingerprint()
    assert fp1 != fp0
    child = m.pose_estimator["body"]
    child.load_state_dict(child.state_dict())                          # load through a CHILD: copy_ bumps the versions
    assert m._fingerprint() != fp1
    assert weights_fingerprint(child) == weights_fingerprint(child)


def test_skeleton_from_metadata_follows_the_dataset_lists():
    from pafuse_b200.h3wb import H3WBSkeleton
    sk = H3WBSkeleton()
    same = H3WBSkeleton.from_metadata({"left_side": sk.metadata["left_side"], "right_side": sk.metadata["right_side"]})
    assert same.symmetry_matches_builtin and same.joints_left == sk.joints_left and same.joints_right == sk.joints_right
    swapped = H3WBSkeleton.from_metadata({"left_side": sk.metadata["right_side"], "right_side": sk.metadata["left_side"]})
    assert not swapped.symmetry_matches_builtin and swapped.joints_left == sk.joints_right


# ------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


def _small_model(depth=2, H=3, K=2, seed=1):
    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(synthetic.default_args(depth=depth), sk.joints_left, sk.joints_right, sk, is_train=False,
                         num_proposals=H, sampling_timesteps=K)
    m.load_state_dict(synthetic.synthetic_state_dict(seed=seed, depth=depth), strict=False)
    return m.cuda().eval(), sk


@gpu
def test_randn_kernel_matches_the_numpy_definition_and_any_slicing():
    from pafuse_b200 import utils
    ctx = utils._post_context(torch.device("cuda", 0))
    n = 27 * 134 * 3 * 5
    full = ctx.randn(77, 2, 0, 1, 4 * n, 4 * n).reshape(-1)
    ref = torch.from_numpy(philox_ref.randn(77, 2, 0, 4 * n))
    d = (full.cpu() - ref).abs()
    assert d.max().item() <= 2e-6 and (d == 0).float().mean().item() > 0.999     # fp64 transform, one fp32 rounding
    # clip shard: one contiguous run at an odd offset
    part = ctx.randn(77, 2, n + 1, 1, n, n).reshape(-1)
    assert torch.equal(part, full[n + 1: 2 * n + 1])
    # hypothesis shard: rows of a (B=4, H=5, per) tensor, hypotheses 1..3
    per = n // 5
    shard = ctx.randn(77, 2, per, 4, 2 * per, 5 * per)
    assert torch.equal(shard, full.reshape(4, 5, per)[:, 1:3].reshape(4, 2 * per))
    assert not torch.equal(ctx.randn(78, 2, 0, 1, n, n), full[:n].reshape(1, n))
    assert ctx.randn(77, 2, 0, 0, n, n).numel() == 0


@gpu
def test_philox_sharding_is_bit_identical_to_one_rank_for_both_modes():
    from pafuse_b200 import distributed as pd
    from pafuse_b200 import synthetic
    m, sk = _small_model()
    eng = pd.CudaEngine(m, sk)
    B, H = 4, 3
    x2d, x2df = (t.cuda() for t in synthetic.synthetic_inputs(B, seed=6))
    traj, cam = synthetic.synthetic_trajectory(B, seed=6).cuda(), synthetic.h36m_cam0_intrinsics().cuda()
    one = pd.lift_sharded(eng, x2d, x2df, traj, cam, H, mode="clips", seed=3, rank=0, world=1)
    parts = [pd.lift_sharded(eng, x2d, x2df, traj, cam, H, mode="clips", seed=3, rank=r, world=3, gather=False)
             for r in range(3)]
    assert torch.equal(torch.cat([p.jagg for p in parts]), one.jagg)
    assert torch.equal(torch.cat([p.pagg for p in parts]), one.pagg)
    assert torch.equal(torch.cat([p.select for p in parts]), one.select)
    # hypothesis shards: the local sampler outputs are the columns of the one-rank sampler output
    full = eng.sample(x2d, x2df, H, eng.noise(3, B, H, (0, B), (0, H), x2d.device))
    for r in range(3):
        h0, h1 = pd.shard_range(H, 3, r)
        local = eng.sample(x2d, x2df, h1 - h0, eng.noise(3, B, H, (0, B), (h0, h1), x2d.device))
        assert torch.equal(local, full[:, :, h0:h1])
    other = pd.lift_sharded(eng, x2d, x2df, traj, cam, H, mode="clips", seed=4, rank=0, world=1)
    assert not torch.equal(other.pagg, one.pagg)


@gpu
def test_small_passes_replay_from_a_cuda_graph_bit_identically():
    from pafuse_b200 import synthetic
    m, sk = _small_model(depth=2, H=2, K=3)
    B = 2
    x2d, x2df = (t.cuda() for t in synthetic.synthetic_inputs(B, seed=2))
    noises = [n.cuda() for n in synthetic.synthetic_noise(B, 2, 3, seed=2)]
    m.noise_source = lambda k, shape, device: noises[k]
    ctx = m.native_context()
    ctx.set_graph_max_seqs(0)
    direct = m(x2d, None, input_2d_flip=x2df).clone()
    assert ctx.graph_replays() == 0
    ctx.set_graph_max_seqs(96)
    outs = [m(x2d, None, input_2d_flip=x2df).clone() for _ in range(4)]
    torch.cuda.synchronize()
    assert ctx.graph_replays() > 0                                    # third and later forwards replay (same pointers)
    for o in outs:
        assert torch.equal(o, direct)
    # a different batch size is a different key: still correct
    x2, x2f = (t.cuda() for t in synthetic.synthetic_inputs(1, seed=2))
    n1 = [n[:1].contiguous() for n in noises]
    m.noise_source = lambda k, shape, device: n1[k]
    a = m(x2, None, input_2d_flip=x2f).clone()
    assert torch.equal(a, direct[:1])


@gpu
def test_in_place_weight_edits_reach_the_library():
    from pafuse_b200 import synthetic
    m, sk = _small_model()
    x2d, x2df = (t.cuda() for t in synthetic.synthetic_inputs(1, seed=2))
    noises = [n.cuda() for n in synthetic.synthetic_noise(1, 3, 2, seed=2)]
    m.noise_source = lambda k, shape, device: noises[k]
    a = m(x2d, None, input_2d_flip=x2df).clone()
    original = {k: v.clone() for k, v in m.pose_estimator["face"].state_dict().items()}
    with torch.no_grad():
        m.pose_estimator["face"].head[1].bias.add_(0.25)              # no load_state_dict, no .to(): only a version bump
    b = m(x2d, None, input_2d_flip=x2df).clone()
    face = sk.parts_joint_indices["face"]
    assert not torch.equal(a[..., face, :], b[..., face, :])
    body = sk.parts_joint_indices["body"]
    assert torch.equal(a[..., body, :], b[..., body, :])
    m.pose_estimator["face"].load_state_dict(original)                # through the CHILD module
    c = m(x2d, None, input_2d_flip=x2df)
    assert torch.equal(a, c)


@gpu
def test_nan_weights_propagate_and_out_of_range_weights_are_refused():
    from pafuse_b200 import _native, synthetic
    m, sk = _small_model(depth=1, H=1, K=1)
    x2d, x2df = (t.cuda() for t in synthetic.synthetic_inputs(1, seed=2))
    with torch.no_grad():
        m.pose_estimator["hands"].STEblocks[0].mlp.fc1.weight[3, 5] = float("nan")
    out = m(x2d, None, input_2d_flip=x2df)
    hands = sk.parts_joint_indices["left_hand"]
    assert torch.isnan(out[..., hands, :]).any()                       # like fp32 torch: not a finite wrong pose
    assert torch.isfinite(out[..., sk.parts_joint_indices["body"], :]).all()
    with torch.no_grad():
        m.pose_estimator["hands"].STEblocks[0].mlp.fc1.weight[3, 5] = 300.0     # 300 * 2^8 > fp16 max
    with pytest.raises(_native.PafuseError, match="magnitude"):
        m(x2d, None, input_2d_flip=x2df)


@gpu
def test_data_parallel_wrapper_with_kwargs_on_one_device():
    """main_h3wb.py:699-705 wraps the model in nn.DataParallel and calls it with a keyword argument."""
    from pafuse_b200 import synthetic
    m, sk = _small_model()
    x2d, x2df = (t.cuda() for t in synthetic.synthetic_inputs(3, seed=2))
    noises = [n.cuda() for n in synthetic.synthetic_noise(3, 3, 2, seed=2)]
    m.noise_source = lambda k, shape, device: noises[k]
    plain = m(x2d, None, input_2d_flip=x2df).clone()
    dp = torch.nn.DataParallel(m, device_ids=[0])
    assert torch.equal(dp(x2d, None, input_2d_flip=x2df), plain)


@gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_one_process_and_data_parallel_scatter():
    """A context on cuda:1 after cuda:0 was used (per-device kernel attributes), and nn.DataParallel over two GPUs:
    clips scattered on dim 0, kwargs scattered with them, results gathered on cuda:0."""
    import pafuse_b200
    from pafuse_b200 import synthetic
    m, sk = _small_model(depth=2, H=2, K=2)
    B = 4
    x2d, x2df = synthetic.synthetic_inputs(B, seed=2)
    noises = synthetic.synthetic_noise(B, 2, 2, seed=2)
    m.noise_source = lambda k, shape, device: noises[k].to(device)
    on0 = m(x2d.cuda(0), None, input_2d_flip=x2df.cuda(0)).clone()
    m1 = m.to("cuda:1")
    on1 = m1(x2d.cuda(1), None, input_2d_flip=x2df.cuda(1))
    assert on1.device.index == 1 and torch.equal(on1.cpu(), on0.cpu())
    wb = pafuse_b200.wb_pose_from_parts(on1.clone(), sk)
    assert wb.device.index == 1
    m0 = m1.to("cuda:0")
    halves = {0: [n[:2].contiguous() for n in noises], 1: [n[2:].contiguous() for n in noises]}
    m0.noise_source = lambda k, shape, device: halves[device.index][k].to(device)
    dp = torch.nn.DataParallel(m0, device_ids=[0, 1])
    for _ in range(2):                                                # second call: replica contexts are reused
        out = dp(x2d.cuda(0), None, input_2d_flip=x2df.cuda(0))
        assert out.device.index == 0 and torch.equal(out.cpu(), on0.cpu())


@gpu
def test_in_the_wild_writes_the_reference_npy_format(tmp_path):
    from pafuse_b200 import in_the_wild, synthetic
    m, sk = _small_model(depth=1, H=2, K=2)
    T = 40
    g = torch.Generator().manual_seed(3)
    det = torch.rand(T, 133, 3, generator=g) * torch.tensor([1920.0, 1080.0, 1.0])
    kp = in_the_wild.keypoints_from_openpifpaf(det, 1920, 1080)
    torch.manual_seed(0)
    res = in_the_wild.lift_video(m, sk, kp, video_name="clip", out_dir=str(tmp_path))
    path = os.path.join(str(tmp_path), "clip", "test_3d_clip_output.npy")       # h3wb_diffusion.py:136
    assert res["path"] == path and os.path.isfile(path)
    arr = np.load(path, allow_pickle=True)
    assert arr.shape == (2, 2, T, 134, 3) and arr.dtype == np.float32         # (K,H,T,134,3), :121
    assert np.array_equal(arr, res["prediction"].cpu().numpy())


@gpu
def test_reassembly_tables_are_cached_and_follow_a_changed_dataset():
    import pafuse_b200
    from pafuse_b200.h3wb import H3WBSkeleton
    x = torch.randn(5, 134, 3, device="cuda")
    ds = H3WBSkeleton()
    a = pafuse_b200.wb_pose_from_parts(x.clone(), ds)
    b = pafuse_b200.wb_pose_from_parts(x.clone(), ds)                 # second call: resident table, no upload
    assert torch.equal(a, b)
    ds2 = H3WBSkeleton()
    ds2.parts_connection_indices = {"face": 2, "left_hand": 10, "right_hand": 11}
    c = pafuse_b200.wb_pose_from_parts(x.clone(), ds2)
    face = ds.parts_joint_indices["face"]
    assert not torch.equal(a[:, face], c[:, face])
    assert torch.equal(pafuse_b200.wb_pose_from_parts(x.clone(), ds), a)


@pytest.mark.gpu
@pytest.mark.parametrize("switch", ["PAFUSE_ATT_PIPE=0", "PAFUSE_ATT_STAGES=3", "PAFUSE_ATT_SEP=0", "PAFUSE_ATT_NACC=2"])
def test_opt_in_attention_variants_stay_parity_green(switch):
    """The attention switches are read once per process, so every variant runs the attention parity cases in a child
    process: the order without the second output buffer, three units in flight, aliased layout only, two PV accumulators."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    name, value = switch.split("=")
    env = dict(os.environ, **{name: value})
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_attention or test_qkv_gemm_head_plane_epilogue_feeds_attention", "-p", "no:cacheprovider"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]
