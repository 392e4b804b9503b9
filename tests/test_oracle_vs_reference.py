"""CPU, build container only: the oracle against the UNMODIFIED reference imported live from
/root/reference on inputs that are NOT in the frozen fixtures (skipped where the reference is absent,
e.g. on the GPU box)."""
import pytest
import torch

from oracle import pafuse_oracle as orc
from oracle import ref_harness
from pafuse_b200 import synthetic
from pafuse_b200.h3wb import H3WBSkeleton, merged_part_indices

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("B,H,K,depth,seed", [(1, 2, 2, 1, 7), (2, 1, 3, 2, 11)])
def test_sampler_live(B, H, K, depth, seed):
    sk = H3WBSkeleton()
    args = synthetic.default_args(depth=depth)
    sd = synthetic.synthetic_state_dict(seed=seed, depth=depth)
    x2d, x2df = synthetic.synthetic_inputs(B, seed=seed)
    noises = synthetic.synthetic_noise(B, H, K, seed=seed)
    model, _ = ref_harness.build_reference_model(args, H3WBSkeleton(), sd, H, K)
    ref = ref_harness.reference_forward(model, x2d, x2df, noises)
    out = orc.ddim_sample_flip(sd, merged_part_indices(sk.parts_joint_indices), x2d, x2df, noises, sk.joints_left,
                               sk.joints_right, H, K, depth=depth)
    assert torch.equal(out, ref)


def test_single_denoiser_live():
    ref = ref_harness.import_reference()
    torch.manual_seed(3)
    J, C, F, depth = 24, 384, 27, 1
    net = ref.mixste.MixSTE2(num_frame=F, num_joints=J, in_chans=5, embed_dim_ratio=C, depth=depth, num_heads=8,
                             mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_path_rate=0, is_train=False).eval()
    with torch.no_grad():
        net.Spatial_pos_embed.normal_(std=0.02)
        net.Temporal_pos_embed.normal_(std=0.02)
        x2d, x3d = torch.rand(2, F, J, 2) * 2 - 1, torch.randn(2, 3, F, J, 3)
        t = torch.full((2,), 599, dtype=torch.long)
        want = net(x2d, x3d, t)
    got = orc.mixste_forward({k: v for k, v in net.state_dict().items()}, x2d, x3d, t, depth=depth)
    assert torch.equal(got, want)


def test_post_processing_live():
    ref = ref_harness.import_reference()
    torch.manual_seed(5)
    pose = torch.randn(2, 3, 2, 27, 134, 3)
    sk = H3WBSkeleton()
    a = pose.clone()
    want = ref.utils.wb_pose_from_parts(a, H3WBSkeleton())
    got, after = orc.wb_pose_from_parts(pose, sk.parts_joint_indices, sk.parts_connection_indices)
    assert torch.equal(got, want) and torch.equal(after, a)
    gt = torch.randn(4, 27, 134, 3)
    assert torch.equal(orc.center_pose_parts(gt, sk.parts_joint_indices, sk.root_indices),
                       ref.utils.center_pose_parts(gt.clone(), H3WBSkeleton()))
    X = torch.randn(6, 134, 3)
    X[..., 2] += 4.0
    cam = synthetic.h36m_cam0_intrinsics().repeat(6, 1)
    assert torch.equal(orc.project_to_2d(X, cam), ref.camera.project_to_2d(X, cam))


def test_reference_known_answer_function_runs():
    ref = ref_harness.import_reference()
    ds = H3WBSkeleton()
    ds.root_indices = {"body": 0, "face": 1, "left_hand": 10, "right_hand": 11}
    ref.utils.test_funcs(ds)                                          # passes only with roots == connection joints
    with pytest.raises(AssertionError):
        ref.utils.test_funcs(H3WBSkeleton())                          # stale with the shipped roots
