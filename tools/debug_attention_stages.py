"""Attention kernel at unit counts that make every CTA re-use its ring slots many times (S = 40 sequences: 18-37 units
per CTA), against fp64 torch -- the case that exposed the slot-sharing hazard of the three-stage variant
(PAFUSE_ATT_STAGES=3).

    CUDA_LAUNCH_BLOCKING=1 [PAFUSE_ATT_STAGES=3] python tools/debug_attention_stages.py
"""
import sys, torch
sys.path.insert(0, '.')
from pafuse_b200 import _native
c = _native.NativeContext(27, 134, 1, 8, [32], [[0]], list(range(134)), 1.0, 1, torch.device("cuda", 0))
for (J, C, temporal, S) in [(68, 224, True, 10), (68, 224, True, 40), (42, 256, False, 40), (42, 256, True, 40), (68, 224, False, 40), (24, 384, True, 40)]:
    torch.manual_seed(1)
    F, hd = 27, C // 8
    qkv = torch.randn(S * F * J, 3 * C, device="cuda")
    try:
        out = c.attention(qkv, S, J, C, temporal)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAIL", J, C, temporal, S, str(e)[:200]); break
    t = qkv.double().reshape(S, F, J, 3, 8, hd)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    perm = (0, 2, 3, 1, 4) if temporal else (0, 1, 3, 2, 4)
    q, k, v = (z.permute(*perm) for z in (q, k, v))
    a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v
    a = a.permute(0, 3, 1, 2, 4) if temporal else a.permute(0, 1, 3, 2, 4)
    print("ok", J, C, temporal, S, (out.double() - a.reshape(S * F * J, C)).abs().max().item())
