"""CPU: host-side mirror of the reference interface (constructor, state_dict, schedule, tables)."""
import torch

import pafuse_b200
from oracle import pafuse_oracle as orc
from pafuse_b200 import synthetic
from pafuse_b200.h3wb import H3WBSkeleton, flip_permutation
from pafuse_b200.mixste import sinusoidal_embedding_cpu
from pafuse_b200.utils import connection_table
from pafuse_testlib import load_golden


def make_model(depth=8, H=5, K=5, flip=True):
    sk = H3WBSkeleton()
    args = synthetic.default_args(depth=depth, test_time_augmentation=flip)
    return pafuse_b200.D3DP(args, sk.joints_left, sk.joints_right, sk, is_train=False, num_proposals=H,
                            sampling_timesteps=K), sk


def test_state_dict_layout_matches_reference_checkpoint():
    m, _ = make_model()
    sd = m.state_dict()
    assert len(sd) == 636                                             # 12 fp64 buffers + 3 x 208 tensors
    buffers = [k for k in sd if not k.startswith("pose_estimator.")]
    assert len(buffers) == 12 and all(sd[k].dtype == torch.float64 and sd[k].shape == (1000,) for k in buffers)
    counts = {p: sum(v.numel() for v in m.pose_estimator[p].parameters()) for p in ("body", "face", "hands")}
    assert counts == {"body": 19558275, "face": 6687971, "hands": 8718083}
    assert sd["pose_estimator.body.STEblocks.3.attn.qkv.weight"].shape == (1152, 384)
    assert sd["pose_estimator.face.Spatial_pos_embed"].shape == (1, 68, 224)
    assert sd["pose_estimator.hands.head.1.weight"].shape == (3, 256)
    assert set(synthetic.synthetic_state_dict(depth=8)) == {k for k in sd if k.startswith("pose_estimator.")}


def test_schedule_buffers_bit_identical_to_reference():
    m, _ = make_model()
    g = load_golden("schedule")
    for k in ("alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
        assert (getattr(m, k).numpy() == g[k]).all()


def test_loads_dataparallel_checkpoints():
    m, _ = make_model(depth=2)
    sd = {"module." + k: v for k, v in synthetic.synthetic_state_dict(depth=2).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith("pose_estimator") for k in missing)
    w = m.state_dict()["pose_estimator.face.TTEblocks.1.mlp.fc2.weight"]
    assert torch.equal(w, sd["module.pose_estimator.face.TTEblocks.1.mlp.fc2.weight"])


def test_time_pairs_and_coefficients_match_oracle():
    for K in (1, 5, 10):
        m, _ = make_model(K=K)
        pairs = m.sampling_time_pairs()
        assert pairs == orc.sampling_times(1000, K)
        ac = orc.cosine_alphas_cumprod(1000)
        for t, tn in pairs:
            sr, srm1, san, c, sigma = m.step_coefficients(t, tn)
            o = orc.ddim_coefficients(ac, t, tn)
            assert sr == float(o[0]) and srm1 == float(o[1])
            if tn >= 0:
                assert (san, c, sigma) == (float(o[2]), float(o[3]), float(o[4]))


def test_sinusoidal_embedding_matches_oracle():
    for C in (384, 224, 256):
        for t in (999, 799, 199, 0):
            a = sinusoidal_embedding_cpu(t, C)
            b = orc.sinusoidal_embedding(torch.tensor([float(t)]), C).reshape(-1)
            assert torch.equal(a, b)


def test_flip_permutation_is_the_reference_index_swap():
    sk = H3WBSkeleton()
    perm = flip_permutation(sk.joints_left, sk.joints_right)
    x = torch.arange(134.0)[:, None].expand(134, 3).clone()
    ref = x.clone()
    ref[sk.joints_left + sk.joints_right, :] = ref[sk.joints_right + sk.joints_left, :]   # diffusionpose.py:197-198
    assert torch.equal(x[perm], ref)
    assert [perm[perm[j]] for j in range(134)] == list(range(134))                         # involution


def test_connection_table():
    conn = connection_table(H3WBSkeleton(), 134)
    assert conn[:24] == [0] * 24 and set(conn[24:92]) == {1} and set(conn[92:113]) == {10} and set(conn[113:]) == {11}


def test_training_branch_is_refused():
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(synthetic.default_args(depth=1), sk.joints_left, sk.joints_right, sk, is_train=True)
    x2d, x2df = synthetic.synthetic_inputs(1)
    try:
        m(x2d, None, input_2d_flip=x2df)
    except NotImplementedError:
        return
    raise AssertionError("training forward must raise")


def test_config_loader_reads_the_reference_yaml_layout(tmp_path):
    from pafuse_b200 import config
    cfg = config.load_config()
    assert cfg.model.number_of_frames == 27 and cfg.ft2d.num_proposals == 10 and cfg.general.part_based_model is True
    y = tmp_path / "config.yaml"
    y.write_text("general:\n  part_based_model: True\nmodel:\n  dep: 4  # depth\n  number_of_frames: 27\nft2d:\n  sampling_timesteps: 3\n")
    cfg = config.load_config(str(y), overrides=["ft2d.num_proposals=20", "model.test_time_augmentation=false"])
    assert cfg.model.dep == 4 and cfg.ft2d.sampling_timesteps == 3 and cfg.ft2d.num_proposals == 20
    assert cfg.model.test_time_augmentation is False and cfg.data.num_kps == 134      # default kept


def test_checkpoint_file_in_the_reference_format_loads(tmp_path):
    """save_state layout of common/logging.py:94-104 with DataParallel key prefixes."""
    import pafuse_b200
    from pafuse_b200 import config, synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    sd = synthetic.synthetic_state_dict(seed=2, depth=1)
    path = tmp_path / "pafuse_model.bin"
    torch.save({"epoch": 3, "lr": 1e-4, "optimizer": {}, "model_pos": {"module." + k: v for k, v in sd.items()}}, path)
    loaded = config.load_checkpoint(str(path))
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(synthetic.default_args(depth=1), sk.joints_left, sk.joints_right, sk, is_train=False)
    missing, unexpected = m.load_state_dict(loaded, strict=False)
    assert not unexpected and all(not k.startswith("pose_estimator") for k in missing)
    k = "pose_estimator.face.STEblocks.0.attn.qkv.weight"
    assert torch.equal(m.state_dict()[k], sd[k])
