"""Measured deviation of the sm_100a path from the reference's golden outputs (GPU box).

    python tools/parity_report.py > profiles/r1_parity.txt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import pafuse_b200
    from pafuse_b200.h3wb import H3WBSkeleton
    from pafuse_testlib import build_case
    sk = H3WBSkeleton()
    print("case                      max|d| (m)   max |d|/max(|ref|,1e-2)   MPJPE delta (mm)   fraction of tolerance used")
    for name in ("cfg1_B2_H1_K1", "tiny_B1_H3_K2_d2", "noflip_B2_H1_K2_d2", "small_B2_H2_K3"):
        c = build_case(name)
        m = pafuse_b200.D3DP(c["args"], sk.joints_left, sk.joints_right, sk, is_train=False, num_proposals=c["H"],
                             sampling_timesteps=c["K"])
        m.load_state_dict(c["sd"], strict=False)
        m = m.cuda().eval()
        noises = c["noises"]
        m.noise_source = lambda k, shape, device: noises[k].to(device)
        out = m(c["x2d"].cuda(), None, input_2d_flip=c["x2df"].cuda() if c["flip"] else None).double().cpu()
        ref = c["golden"]["out"].double()
        d = (out - ref).abs()
        rel = (d / ref.abs().clamp_min(1e-2)).max().item()
        used = (d / (1e-3 * ref.abs() + 2e-5)).max().item()
        mpjpe = (out - ref).norm(dim=-1).mean().item() * 1e3
        print(f"{name:24s}  {d.max().item():.3e}    {rel:.3e}                 {mpjpe:.6f}           {used:.3f}")


if __name__ == "__main__":
    main()
