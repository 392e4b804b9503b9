"""Caller-side rows: flip-TTA input construction, eval_data_prepare tiling, detection ->
keypoints normalisation, clip stitching, and the in-the-wild driver built from them.

CPU tests pin the oracle against vectors the reference itself produced (tests/golden/make_golden_callers.py);
GPU tests compare the kernels (through the C ABI) with the oracle bit for bit.
"""
import numpy as np
import pytest
import torch

from oracle import pafuse_oracle as orc
from pafuse_testlib import load_golden


def _golden():
    g = load_golden("callers")
    return {k: torch.from_numpy(g[k]) for k in g.files}


# ------------------------------------------------------------------ oracle against the reference's vectors (CPU)
def test_oracle_keypoints_match_reference():
    g = _golden()
    w, h = (int(v) for v in g["wh"])
    assert torch.equal(orc.keypoints_from_openpifpaf(g["det"], w, h), g["kp"])


@pytest.mark.parametrize("T", [5, 27, 54, 70])
def test_oracle_prepare_matches_reference(T, skeleton):
    g = _golden()
    seq = g["kp"][:T]
    assert torch.equal(orc.eval_data_prepare(27, seq), g[f"clips_T{T}"])
    flip = orc.flip_inputs_2d(seq, skeleton.kps_left(), skeleton.kps_right())
    assert torch.equal(orc.eval_data_prepare(27, flip), g[f"clips_flip_T{T}"])


def test_oracle_flip_is_an_involution(skeleton):
    x = torch.randn(3, 27, 134, 2)
    L, R = skeleton.kps_left(), skeleton.kps_right()
    assert torch.equal(orc.flip_inputs_2d(orc.flip_inputs_2d(x, L, R), L, R), x)


# ------------------------------------------------------------------ kernels against the oracle (GPU)
@pytest.mark.gpu
def test_keypoints_kernel_bit_exact():
    from pafuse_b200 import in_the_wild as itw
    g = _golden()
    w, h = (int(v) for v in g["wh"])
    got = itw.keypoints_from_openpifpaf(g["det"], w, h)
    assert torch.equal(got.cpu(), g["kp"])


@pytest.mark.gpu
@pytest.mark.parametrize("T", [1, 5, 26, 27, 28, 54, 70, 3000])
def test_prepare_and_stitch_kernels_bit_exact(T, skeleton):
    import pafuse_b200
    from pafuse_b200 import in_the_wild as itw
    gen = torch.Generator().manual_seed(T)
    seq = torch.rand(T, 134, 2, generator=gen) * 2 - 1
    L, R = skeleton.kps_left(), skeleton.kps_right()
    clips, flip = pafuse_b200.eval_data_prepare(27, seq.cuda()[None], L, R)
    assert torch.equal(clips.cpu(), orc.eval_data_prepare(27, seq))
    assert torch.equal(flip.cpu(), orc.eval_data_prepare(27, orc.flip_inputs_2d(seq, L, R)))
    only = pafuse_b200.eval_data_prepare(27, seq.cuda())
    assert torch.equal(only, clips)
    n = clips.shape[0]
    pred = torch.randn(n, 2, 3, 27, 134, 3, generator=gen)
    assert torch.equal(itw.stitch_predictions(pred.cuda(), T).cpu(), orc.stitch_clips(pred, T))


@pytest.mark.gpu
def test_golden_clips_from_the_reference(skeleton):
    import pafuse_b200
    g = _golden()
    for T in (5, 27, 54, 70):
        clips, flip = pafuse_b200.eval_data_prepare(27, g["kp"][:T].cuda(), skeleton.kps_left(), skeleton.kps_right())
        assert torch.equal(clips.cpu(), g[f"clips_T{T}"]) and torch.equal(flip.cpu(), g[f"clips_flip_T{T}"])


@pytest.mark.gpu
def test_in_the_wild_driver_against_oracle(skeleton):
    """BASELINE config 5 in miniature: 70 synthetic detection frames -> 3 clips -> lift -> stitch -> mean pose."""
    import pafuse_b200
    from pafuse_b200 import in_the_wild as itw
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton, merged_part_indices
    g = _golden()
    w, h = (int(v) for v in g["wh"])
    T, H, K, depth = 70, 2, 2, 2
    sd = synthetic.synthetic_state_dict(seed=3, depth=depth)
    model = pafuse_b200.D3DP(synthetic.default_args(depth=depth), skeleton.joints_left, skeleton.joints_right, skeleton,
                             is_train=False, num_proposals=H, sampling_timesteps=K)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    noises = synthetic.synthetic_noise(3, H, K, seed=9)
    model.noise_source = lambda k, shape, device: noises[k].to(device)
    kp = itw.keypoints_from_openpifpaf(g["det"], w, h)
    res = itw.lift_video(model, H3WBSkeleton(), kp, receptive_field=27, bs=8)
    assert res["prediction"].shape == (K, H, T, 134, 3) and res["mean_pose"].shape == (T, 134, 3)

    L, R = skeleton.kps_left(), skeleton.kps_right()
    seq = orc.keypoints_from_openpifpaf(g["det"], w, h)
    x2d = orc.eval_data_prepare(27, seq)
    x2df = orc.eval_data_prepare(27, orc.flip_inputs_2d(seq, L, R))
    parts = merged_part_indices(skeleton.parts_joint_indices)
    ref = orc.ddim_sample_flip(sd, parts, x2d, x2df, noises, skeleton.joints_left, skeleton.joints_right, H, K, depth=depth)
    ref, _ = orc.wb_pose_from_parts(ref, skeleton.parts_joint_indices, skeleton.parts_connection_indices)
    ref = orc.stitch_clips(ref, T)
    d = (res["prediction"].cpu().double() - ref.double()).abs()
    assert not (d > 1e-3 * ref.double().abs() + 2e-5).any(), d.max().item()
    assert torch.allclose(res["mean_pose"].cpu(), ref[-1].mean(dim=0), rtol=1e-3, atol=2e-5)


# ------------------------------------------------------------------ GT-dependent metrics
def test_oracle_metrics_match_reference_loss_functions():
    g = _golden()
    pred, target, t2d, rep = g["m_pred"], g["m_target"], g["m_target_2d"], g["m_reproj"]
    assert torch.allclose(orc.mpjpe_j_best(pred, target), g["m_j_best"], rtol=1e-6, atol=0)
    assert torch.allclose(orc.mpjpe_p_agg(pred, target), g["m_p_agg"], rtol=1e-6, atol=0)
    assert torch.allclose(orc.mpjpe_j_agg(pred, target, rep, t2d), g["m_j_agg"], rtol=1e-6, atol=0)
    assert torch.allclose(orc.mpjpe_p_best(pred, target), g["m_p_best"], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_metrics_kernel_against_reference_values():
    from pafuse_b200 import loss
    g = _golden()
    pred, target, t2d, rep = (g[k].cuda() for k in ("m_pred", "m_target", "m_target_2d", "m_reproj"))
    tol = dict(rtol=2e-6, atol=0)                                      # fp32 means of the reference vs fp64 sums here
    assert torch.allclose(loss.mpjpe_diffusion_all_min(pred, target).cpu(), g["m_j_best"], **tol)
    assert torch.allclose(loss.mpjpe_diffusion_all_min(pred, target, mean_pos=True).cpu(), g["m_p_agg"], **tol)
    assert torch.allclose(loss.mpjpe_diffusion_reproj(pred, target, rep, t2d).cpu(), g["m_j_agg"], **tol)
    assert torch.allclose(loss.mpjpe_diffusion(pred, target)[0].cpu(), g["m_p_best"], **tol)


@pytest.mark.gpu
def test_evaluate_metrics_is_consistent_with_aggregation():
    """J-Agg / P-Agg errors of evaluate_metrics == errors of the poses aggregate_hypotheses returns."""
    import pafuse_b200
    from pafuse_b200 import loss, synthetic
    g = _golden()
    pred, target, t2d = (g[k].cuda() for k in ("m_pred", "m_target", "m_target_2d"))
    B = pred.shape[0]
    traj, cam = synthetic.synthetic_trajectory(B, seed=2).cuda(), synthetic.h36m_cam0_intrinsics().cuda()
    m = loss.evaluate_metrics(pred, target, traj, cam, t2d)
    jagg, pagg = pafuse_b200.aggregate_hypotheses(pred, traj, cam, t2d)
    ej = (jagg - target[:, None]).norm(dim=-1).permute(1, 0, 2, 3).reshape(pred.shape[1], -1).double().mean(dim=-1)
    ep = (pagg - target[:, None]).norm(dim=-1).permute(1, 0, 2, 3).reshape(pred.shape[1], -1).double().mean(dim=-1)
    assert torch.allclose(m["J-Agg"].double(), ej, rtol=2e-6) and torch.allclose(m["P-Agg"].double(), ep, rtol=2e-6)
    assert (m["J-Best"] <= m["J-Agg"] + 1e-7).all() and (m["J-Best"] <= m["P-Best"] + 1e-7).all()


# ------------------------------------------------------------------ part-based protocols + evaluate() accumulation
def _golden_parts():
    import os

    import numpy as np
    from pafuse_testlib import GOLDEN
    g = np.load(os.path.join(GOLDEN, "metrics_parts.npz"))
    return {k: torch.from_numpy(g[k]) for k in g.files}


def test_oracle_part_based_metrics_match_reference(skeleton):
    """common/loss.py with part_based=True (values frozen by tests/golden/make_golden_metrics_parts.py)."""
    g, gp = _golden(), _golden_parts()
    pred, target = g["m_pred"], g["m_target"]
    best, parts = orc.mpjpe_p_best_parts(pred, target, skeleton.parts_joint_indices, skeleton.root_indices)
    assert torch.allclose(best, gp["p_best_pb"], rtol=1e-6, atol=0)
    agg, agg_parts = orc.mpjpe_p_agg_parts(pred, target, skeleton.parts_joint_indices, skeleton.root_indices)
    assert torch.allclose(agg, gp["p_agg_pb"], rtol=1e-6, atol=0)
    for n in skeleton.parts_joint_indices:
        assert torch.allclose(parts[n], gp[f"p_best_pb_{n}"], rtol=1e-6, atol=0)
        assert torch.allclose(agg_parts[n], gp[f"p_agg_pb_{n}"], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_part_based_metrics_kernel_against_reference_values(skeleton):
    from pafuse_b200 import loss
    g, gp = _golden(), _golden_parts()
    pred, target = g["m_pred"].cuda(), g["m_target"].cuda()
    tol = dict(rtol=2e-6, atol=0)
    best, parts = loss.mpjpe_diffusion(pred, target, part_based=True, dataset=skeleton)
    agg, agg_parts = loss.mpjpe_diffusion_all_min(pred, target, mean_pos=True, part_based=True, dataset=skeleton)
    assert torch.allclose(best.cpu(), gp["p_best_pb"], **tol) and torch.allclose(agg.cpu(), gp["p_agg_pb"], **tol)
    assert set(parts) == set(agg_parts) == {"body", "face", "left_hand", "right_hand"}
    for n in parts:
        assert torch.allclose(parts[n].cpu(), gp[f"p_best_pb_{n}"], **tol), n
        assert torch.allclose(agg_parts[n].cpu(), gp[f"p_agg_pb_{n}"], **tol), n


@pytest.mark.gpu
def test_evaluator_reproduces_the_reference_accumulation(skeleton):
    """Evaluator.update per sub-batch == the epoch sums / N * 1000 of evaluate() (main_h3wb.py:364-379,417-432)."""
    from pafuse_b200 import loss
    g, gp = _golden(), _golden_parts()
    pred, target, t2d, rep = (g[k].cuda() for k in ("m_pred", "m_target", "m_target_2d", "m_reproj"))
    ev = loss.Evaluator(skeleton, pred.shape[1])
    for b in range(pred.shape[0]):
        # J-Agg from the frozen reprojection: same path as evaluate_metrics, reproj given instead of cam + traj
        p, t = pred[b:b + 1], target[b:b + 1]
        m = loss._means(p, t, x2d=t2d[b:b + 1], reproj=rep[b:b + 1])
        ev_metrics = {"J-Best": m[:, 0].float(), "P-Agg": m[:, 1].float(), "J-Agg": m[:, 2].float(),
                      "P-Best": m[:, 3:].min(dim=1).values.float()}
        orig = loss.evaluate_metrics
        loss.evaluate_metrics = lambda *a, **k: ev_metrics
        try:
            ev.update(p, t, None, None, t2d[b:b + 1])
        finally:
            loss.evaluate_metrics = orig
    res = ev.results()
    for k in res:
        assert torch.allclose(res[k].cpu(), gp["eval_" + k], rtol=3e-6, atol=0), k
    lines = ev.log_lines(action="Walking")
    assert lines[0] == "----Walking----" and lines[-1] == "----------"
    assert lines[1] == "step 0 : Protocol #1 Error (MPJPE) J_Best: %f mm" % res["J_Best"][0].item()
    assert sum("Part-Based HANDS" in l for l in lines) == 2 * pred.shape[1]


@pytest.mark.gpu
def test_metrics_of_an_empty_batch_are_nan_not_a_crash(skeleton):
    from pafuse_b200 import loss
    pred = torch.zeros(0, 2, 3, 27, 134, 3, device="cuda")
    target = torch.zeros(0, 27, 134, 3, device="cuda")
    assert torch.isnan(loss.mpjpe_diffusion_all_min(pred, target)).all()
    best, parts = loss.mpjpe_diffusion(pred, target, part_based=True, dataset=skeleton)
    assert torch.isnan(best).all() and set(parts) == {"body", "face", "left_hand", "right_hand"}
