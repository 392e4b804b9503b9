"""One denoiser pass (pred_parts) at the bench's per-chunk shape, for ncu.

    ncu ... python tools/profile_pass.py [seqs] [passes]

seqs = clip x hypothesis sequences in the pass (default 256 = one workspace chunk of the bench
workload); every pass launches 3 time-MLPs + 3 x (embed + 16 x (2 LayerNorm + 4 GEMM + attention) + head)
= 345 kernels, in part order body, face, hands.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    H = 4
    B = seqs // H
    sk = H3WBSkeleton()
    model = pafuse_b200.D3DP(synthetic.default_args(depth=8), sk.joints_left, sk.joints_right, sk, is_train=False,
                             num_proposals=H, sampling_timesteps=1)
    model.load_state_dict(synthetic.synthetic_state_dict(seed=1, depth=8), strict=False)
    model = model.cuda().eval()
    x2d, _ = synthetic.synthetic_inputs(B, seed=1)
    x3d = torch.randn(B, H, 27, 134, 3, device="cuda")
    t = torch.full((B,), 999, dtype=torch.long, device="cuda")
    for _ in range(passes):
        out = model.pred_parts(x2d.cuda(), x3d, t)
    torch.cuda.synchronize()
    print("pass ok", tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
