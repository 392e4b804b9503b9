"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run once in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Writes small ``.npz`` files next to this script.  Inputs, weights and noise are
NOT stored: they are regenerated from seeds by ``pafuse_b200.synthetic`` (CPU
generators, name-keyed), only the reference OUTPUTS are frozen.  A sha256 of
the regenerated inputs is stored so a drifting generator is detected.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from pafuse_b200 import synthetic  # noqa: E402
from pafuse_b200.h3wb import H3WBSkeleton  # noqa: E402

# (name, B, H, K, depth, flip)
CASES = [
    ("cfg1_B2_H1_K1", 2, 1, 1, 8, True),          # BASELINE.json configs[0] (CPU plumbing)
    ("small_B2_H2_K3", 2, 2, 3, 8, True),         # multi-hypothesis, multi-step (exercises the DDIM update + noise draws)
    ("tiny_B1_H3_K2_d2", 1, 3, 2, 2, True),       # shallow model: fast CPU check of every code path
    ("noflip_B2_H1_K2_d2", 2, 1, 2, 2, False),    # non-TTA sampler (valid for H=1 only in the reference)
]


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    sk = H3WBSkeleton()
    for name, B, H, K, depth, flip in CASES:
        args = synthetic.default_args(depth=depth, test_time_augmentation=flip)
        sd = synthetic.synthetic_state_dict(seed=1, depth=depth)
        x2d, x2d_flip = synthetic.synthetic_inputs(B, seed=1)
        noises = synthetic.synthetic_noise(B, H, K, seed=1)
        model, ref = ref_harness.build_reference_model(args, H3WBSkeleton(), sd, H, K)
        out = ref_harness.reference_forward(model, x2d, x2d_flip if flip else None, noises)
        # post-processing with the reference's own functions
        ds = H3WBSkeleton()
        wb_in = out.clone()
        wb = ref.utils.wb_pose_from_parts(wb_in, ds)
        traj = synthetic.synthetic_trajectory(B, seed=1)
        cam = synthetic.h36m_cam0_intrinsics()
        b, k, h, f, j, c = wb.shape
        absd = (wb + traj.unsqueeze(1).unsqueeze(1).repeat(1, k, h, 1, 1, 1)).reshape(b * k * h * f, j, c)
        reproj = ref.camera.project_to_2d(absd, cam.repeat(b * k * h * f, 1)).reshape(b, k, h, f, j, 2)
        # J-Agg pose exactly as common/visualization.py:453-463 (per clip), P-Agg as loss.py:68-70
        tgt = x2d.unsqueeze(1).unsqueeze(1).repeat(1, k, h, 1, 1, 1)
        err2d = torch.norm(reproj - tgt, dim=len(tgt.shape) - 1)
        sel = torch.min(err2d, dim=2, keepdim=True).indices
        jagg = torch.gather(wb, 2, sel.unsqueeze(-1).repeat(1, 1, 1, 1, 1, 3)).squeeze(2)
        pagg = torch.mean(wb, dim=2, keepdim=False)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            out=out.numpy(), wb=wb.numpy(), wb_input_after=wb_in.numpy(), reproj=reproj.numpy(),
            jagg=jagg.numpy(), pagg=pagg.numpy(), select=sel.squeeze(2).numpy(),
            meta=np.array([B, H, K, depth, int(flip)], dtype=np.int64),
            input_digest=np.array(digest(x2d, x2d_flip, *noises, *[sd[k_] for k_ in sorted(sd)][:8])),
        )
        print(name, tuple(out.shape), "ok")

    # schedule known answers straight from the reference buffers
    args = synthetic.default_args()
    model, _ = ref_harness.build_reference_model(args, H3WBSkeleton(), synthetic.synthetic_state_dict(depth=8), 1, 1)
    np.savez_compressed(os.path.join(HERE, "schedule.npz"),
                        alphas_cumprod=model.alphas_cumprod.numpy(),
                        sqrt_recip_alphas_cumprod=model.sqrt_recip_alphas_cumprod.numpy(),
                        sqrt_recipm1_alphas_cumprod=model.sqrt_recipm1_alphas_cumprod.numpy())

    # the reference's only known-answer test, common/utils.py:129-157, run with the
    # roots it was written for ({0,1,10,11}; it fails with the shipped roots)
    ref = ref_harness.import_reference()
    ds = H3WBSkeleton()
    ds.root_indices = {"body": 0, "face": 1, "left_hand": 10, "right_hand": 11}
    ref.utils.test_funcs(ds)
    print("reference test_funcs passed with roots {0,1,10,11}")


if __name__ == "__main__":
    main()
