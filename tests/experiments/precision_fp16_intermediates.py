"""Precision experiment on the CPU oracle (not a test, not collected): see DESIGN.md section 10."""
import sys, time, torch, torch.nn.functional as Fn
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle import pafuse_oracle as orc
from pafuse_b200 import synthetic
from pafuse_b200.h3wb import H3WBSkeleton, merged_part_indices
torch.set_num_threads(8)
FLAGS = set()
def r16(x, name):
    return x.half().float() if name in FLAGS else x
def _attention(w, pre, x, heads):
    G, L, C = x.shape
    hd = C // heads
    qkv = Fn.linear(x, w[pre + "attn.qkv.weight"], w[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(G, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = r16(qkv[0], 'q'), r16(qkv[1], 'k'), r16(qkv[2], 'v')
    a = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    a = r16(a.softmax(dim=-1), 'p')
    o = r16((a @ v).transpose(1, 2).reshape(G, L, C), 'o')
    return Fn.linear(o, w[pre + "attn.proj.weight"], w[pre + "attn.proj.bias"])
def _block(w, pre, x, heads):
    C = x.shape[-1]
    x = x + _attention(w, pre, r16(Fn.layer_norm(x, (C,), w[pre + "norm1.weight"], w[pre + "norm1.bias"], 1e-6), 'a1'), heads)
    h = r16(Fn.layer_norm(x, (C,), w[pre + "norm2.weight"], w[pre + "norm2.bias"], 1e-6), 'a2')
    h = r16(Fn.gelu(Fn.linear(h, w[pre + "mlp.fc1.weight"], w[pre + "mlp.fc1.bias"])), 'h')
    return x + Fn.linear(h, w[pre + "mlp.fc2.weight"], w[pre + "mlp.fc2.bias"])
orc._attention = _attention; orc._block = _block
B, H, K, depth = 1, 5, 5, 8
sk = H3WBSkeleton()
sd = synthetic.synthetic_state_dict(seed=1, depth=depth)
x2d, x2df = synthetic.synthetic_inputs(B, seed=1)
noises = synthetic.synthetic_noise(B, H, K, seed=1)
parts = merged_part_indices(sk.parts_joint_indices)
def run(flags):
    global FLAGS
    FLAGS.clear(); FLAGS.update(flags)
    return orc.ddim_sample_flip(sd, parts, x2d, x2df, noises, sk.joints_left, sk.joints_right, H, K, depth=depth)
with torch.no_grad():
    ref = run([])
    for fl in [['q','k','v'], ['q','k'], ['v'], ['p'], ['o'], ['h'], ['a1'], ['a2'], ['q','k','v','p'], ['q','k','v','p','o','h']]:
        out = run(fl)
        d = out - ref
        print(fl, 'mpjpe mm %.4f' % (d.norm(dim=-1).mean().item()*1e3), 'maxrel %.2e' % ((d.abs()/ref.abs().clamp_min(1e-2)).max().item()), 'maxabs %.2e' % d.abs().max().item(), flush=True)
