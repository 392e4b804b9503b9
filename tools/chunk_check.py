import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pafuse_b200
from pafuse_b200 import synthetic
from pafuse_b200.h3wb import H3WBSkeleton
sk = H3WBSkeleton()
B, H, K = 64, 5, 1
outs = []
for ms in (256, 640):
    m = pafuse_b200.D3DP(synthetic.default_args(depth=8), sk.joints_left, sk.joints_right, sk, is_train=False, num_proposals=H, sampling_timesteps=K)
    m.load_state_dict(synthetic.synthetic_state_dict(seed=1, depth=8), strict=False)
    m.max_seqs = ms
    m = m.cuda().eval()
    noises = synthetic.synthetic_noise(B, H, K, seed=1)
    m.noise_source = lambda k, shape, device: noises[k].to(device)
    x2d, x2df = synthetic.synthetic_inputs(B, seed=1)
    outs.append(m(x2d.cuda(), None, input_2d_flip=x2df.cuda()))
d = (outs[0] - outs[1]).abs().max().item()
print("max |diff| 256 vs 640 chunks:", d, "finite:", bool(torch.isfinite(outs[1]).all()))
