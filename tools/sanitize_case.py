"""Tiny end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box.

    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pafuse_b200
    from pafuse_b200 import in_the_wild as itw
    from pafuse_b200 import loss, synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    depth, H, K, T = 1, 2, 2, 40
    sk = H3WBSkeleton()
    model = pafuse_b200.D3DP(synthetic.default_args(depth=depth), sk.joints_left, sk.joints_right, sk, is_train=False,
                             num_proposals=H, sampling_timesteps=K)
    model.load_state_dict(synthetic.synthetic_state_dict(seed=1, depth=depth), strict=False)
    model = model.cuda().eval()
    det = torch.rand(T, 133, 3) * torch.tensor([1920.0, 1080.0, 1.0])
    kp = itw.keypoints_from_openpifpaf(det, 1920, 1080)
    res = itw.lift_video(model, H3WBSkeleton(), kp, receptive_field=27, bs=8)
    x2d, x2df = pafuse_b200.eval_data_prepare(27, kp, sk.kps_left(), sk.kps_right())
    pred = model(x2d, None, input_2d_flip=x2df)
    wb = pafuse_b200.wb_pose_from_parts(pred.clone(), H3WBSkeleton())
    traj, cam = synthetic.synthetic_trajectory(x2d.shape[0], seed=1).cuda(), synthetic.h36m_cam0_intrinsics().cuda()
    jagg, pagg, sel = pafuse_b200.aggregate_hypotheses(wb, traj, cam, x2d, return_select=True)
    m = loss.evaluate_metrics(wb, jagg[:, -1], traj, cam, x2d)
    # round-2 entry points: part-based metrics, counter-based noise, graph replay of a small pass (second / third forward)
    loss.mpjpe_diffusion(wb, jagg[:, -1], part_based=True, dataset=sk)
    loss.mpjpe_diffusion_all_min(wb, jagg[:, -1], mean_pos=True, part_based=True, dataset=sk)
    model.native_context().randn(1, 0, 3, 2, 1001, 4000)
    noises = [torch.randn(x2d.shape[0], H, 27, 134, 3, device="cuda") for _ in range(K)]
    model.noise_source = lambda k, shape, device: noises[k]
    for _ in range(3):
        pred2 = model(x2d, None, input_2d_flip=x2df)
    torch.cuda.synchronize()
    print("sanitize case ok", tuple(res["prediction"].shape), float(pagg.abs().mean()), {k: float(v[-1]) for k, v in m.items()})


if __name__ == "__main__":
    main()
