"""Precision experiment on the CPU oracle (not a test, not collected): what does dropping the `lo` half of the WEIGHTS of
one GEMM kind cost (two tensor passes A_hi*W_hi + A_lo*W_hi instead of three)?  See DESIGN.md section 9."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle import pafuse_oracle as orc
from pafuse_b200 import synthetic
from pafuse_b200.h3wb import H3WBSkeleton, merged_part_indices
torch.set_num_threads(os.cpu_count() or 1)
B, H, K, depth = 1, 5, 5, 8
sk = H3WBSkeleton()
sd = synthetic.synthetic_state_dict(seed=1, depth=depth)
x2d, x2df = synthetic.synthetic_inputs(B, seed=1)
noises = synthetic.synthetic_noise(B, H, K, seed=1)
parts = merged_part_indices(sk.parts_joint_indices)
def rounded(kinds):
    out = {}
    for k, v in sd.items():
        if any(k.endswith(s + ".weight") for s in kinds) and ("STEblocks" in k or "TTEblocks" in k):
            v = (v * 256).half().float() / 256                      # the hi half of the committed operand (scaled by 2^8)
        out[k] = v
    return out
def run(w):
    return orc.ddim_sample_flip(w, parts, x2d, x2df, noises, sk.joints_left, sk.joints_right, H, K, depth=depth)
with torch.no_grad():
    ref = run(sd)
    for kinds in (["attn.qkv"], ["attn.proj"], ["mlp.fc1"], ["mlp.fc2"], ["attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"]):
        d = run(rounded(kinds)) - ref
        tol = (d.abs() / (1e-3 * ref.abs() + 2e-5)).max().item()
        print(kinds, 'mpjpe mm %.4f' % (d.norm(dim=-1).mean().item() * 1e3), 'maxabs %.2e' % d.abs().max().item(),
              'worst tolerance ratio %.2f' % tol, flush=True)
