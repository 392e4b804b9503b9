"""Shared helpers of the test-suite: golden fixtures and seeded cases."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def build_case(name):
    """Regenerate the seeded inputs of a golden case and check their digest."""
    import hashlib

    import torch

    from pafuse_b200 import synthetic
    g = load_golden(name)
    B, H, K, depth, flip = [int(v) for v in g["meta"]]
    sd = synthetic.synthetic_state_dict(seed=1, depth=depth)
    x2d, x2df = synthetic.synthetic_inputs(B, seed=1)
    noises = synthetic.synthetic_noise(B, H, K, seed=1)
    h = hashlib.sha256()
    for t in (x2d, x2df, *noises, *[sd[k] for k in sorted(sd)][:8]):
        h.update(t.detach().contiguous().numpy().tobytes())
    assert h.hexdigest() == str(g["input_digest"]), "synthetic generators drifted from the golden fixtures"
    keys = ("out", "wb", "wb_input_after", "reproj", "jagg", "pagg", "select", "out_last", "out_frames")
    return dict(g=g, B=B, H=H, K=K, depth=depth, flip=bool(flip), sd=sd, x2d=x2d, x2df=x2df, noises=noises,
                traj=synthetic.synthetic_trajectory(B, seed=1), cam=synthetic.h36m_cam0_intrinsics(),
                args=synthetic.default_args(depth=depth, test_time_augmentation=bool(flip)),
                golden={k: torch.from_numpy(g[k]) for k in keys if k in g.files})
