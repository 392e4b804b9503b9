"""Bisect test_workspace_chunking_is_invisible: full vs chunked workspace under the four
(GEMM, attention) x (tensor-core, CUDA-core) combinations."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def model(c, max_seqs=None, simt_gemm=False, simt_attn=False):
    import pafuse_b200
    from pafuse_b200.h3wb import H3WBSkeleton
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(c["args"], sk.joints_left, sk.joints_right, sk, is_train=False, num_proposals=c["H"],
                         sampling_timesteps=c["K"])
    m.load_state_dict(c["sd"], strict=False)
    m = m.cuda().eval()
    noises = c["noises"]
    m.noise_source = lambda k, shape, device: noises[k].to(device)
    if max_seqs:
        m.max_seqs = max_seqs
        m._native_dirty = True
    ctx = m.native_context()
    ctx.set_debug_simt_gemm(simt_gemm)
    ctx.set_debug_simt_attention(simt_attn)
    return m


def main():
    from pafuse_testlib import build_case
    c = build_case("small_B2_H2_K3")
    x, xf = c["x2d"].cuda(), c["x2df"].cuda()
    for sg in (False, True):
        for sa in (False, True):
            full = model(c, None, sg, sa)(x, None, input_2d_flip=xf)
            full2 = model(c, None, sg, sa)(x, None, input_2d_flip=xf)
            for ms in (3, 4, 1):
                ch = model(c, ms, sg, sa)(x, None, input_2d_flip=xf)
                d = (full - ch).abs()
                nz = d > 0
                joints = sorted(set(torch.nonzero(nz)[:, 4].tolist()))
                steps = sorted(set(torch.nonzero(nz)[:, 1].tolist()))
                print(f"simt_gemm={sg} simt_attn={sa} max_seqs={ms}: rerun-equal={torch.equal(full, full2)} "
                      f"max diff {d.max().item():.3e} n_diff {int(nz.sum())} steps {steps} "
                      f"joints {joints[:6]}..{joints[-3:] if joints else []} ({len(joints)})", flush=True)


if __name__ == "__main__":
    main()
