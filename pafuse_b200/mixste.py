"""Host-side mirror of the reference denoiser module ``MixSTE2``.

Same constructor, parameter names/shapes (so reference ``state_dict``s load
unchanged) and ``forward(x_2d, x_3d, t)`` contract as ``common/mixste.py:141-298``;
the arithmetic runs in the sm_100a library through the C ABI
(``pafuse_pred_parts`` with a one-part table).  Inference branch only: the
training branch of the reference is out of this tier's scope.
"""
from __future__ import annotations

import math
from functools import partial

import torch
from torch import nn

from . import _native


def sinusoidal_embedding_cpu(t: float, dim: int) -> torch.Tensor:
    """SinusoidalPositionEmbeddings (mixste.py:132-139) for one scalar timestep,
    evaluated with CPU torch ops so every rank/device sees the same table."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half) * -e)
    a = torch.tensor([float(t)])[:, None] * f[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1).reshape(-1).float()


class _Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    """Parameter container with the reference Block's sub-module names (mixste.py:84-111)."""

    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _SinusoidalPositionEmbeddings(nn.Module):  # keeps time_mlp indices 1 and 3 for the Linear layers
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


def state_items(module: nn.Module):
    """(sub-key, tensor) pairs of one part denoiser in the C-ABI naming.  Works on ``nn.DataParallel`` replicas too:
    their parameters are plain attributes (``_former_parameters``), which ``state_dict()`` does not list."""
    sd = module.state_dict()
    if sd:
        yield from sd.items()
        return
    for prefix, sub in module.named_modules():
        for k, v in getattr(sub, "_former_parameters", {}).items():
            yield (prefix + "." if prefix else "") + k, v


def weights_fingerprint(module: nn.Module):
    """Cheap identity of the current weights: storage address and in-place version counter of every parameter.
    Changes on ``load_state_dict`` of a child, ``p.data.copy_`` / init functions, optimizer steps, ``.to()``."""
    items = []
    for sub in module.modules():
        for src in (sub._parameters, getattr(sub, "_former_parameters", {})):
            for p in src.values():
                if p is not None:
                    items.append((p.data_ptr(), p._version))
    return hash(tuple(items))


class MixSTE2(nn.Module):
    def __init__(self, num_frame=9, num_joints=17, in_chans=5, embed_dim_ratio=32, depth=4, num_heads=8, mlp_ratio=2.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None,
                 is_train=True):
        super().__init__()
        if num_heads != 8 or mlp_ratio != 2. or in_chans != 5 or not qkv_bias or qk_scale is not None:
            raise NotImplementedError("pafuse_b200.MixSTE2 supports the PAFUSE configuration only "
                                      "(8 heads, mlp_ratio 2, in_chans 5, qkv_bias, default qk scale)")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        C = embed_dim_ratio
        self.is_train = is_train
        self.num_frame, self.num_joints, self.embed_dim, self.block_depth = num_frame, num_joints, C, depth
        self.Spatial_patch_to_embedding = nn.Linear(in_chans, C)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, C))
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, C))
        self.time_mlp = nn.Sequential(_SinusoidalPositionEmbeddings(C), nn.Linear(C, C * 2), nn.GELU(),
                                      nn.Linear(C * 2, C))
        self.STEblocks = nn.ModuleList([_Block(C, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.TTEblocks = nn.ModuleList([_Block(C, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.Spatial_norm = norm_layer(C)
        self.Temporal_norm = norm_layer(C)
        self.head = nn.Sequential(nn.LayerNorm(C), nn.Linear(C, 3))
        self._natives = {}                # device index -> (context, fingerprint of the weights it holds)
        self.max_seqs = 640

    def _native(self, device) -> _native.NativeContext:
        key = torch.device(device).index
        if key is None:
            key = torch.cuda.current_device()
        fp = weights_fingerprint(self)    # the packed fp16 copies follow ANY change of the weights
        held = self._natives.get(key)
        if held is not None and held[1] != fp:
            held[0].close()
            held = None
        if held is None:
            J = self.num_joints
            ctx = _native.NativeContext(self.num_frame, J, self.block_depth, 8, [self.embed_dim], [list(range(J))],
                                        list(range(J)), 1.0, self.max_seqs, torch.device("cuda", key))
            for name, t in state_items(self):
                ctx.set_weight(0, name, t)
            ctx.commit_weights()
            held = self._natives[key] = (ctx, fp)
        return held[0]

    @torch.no_grad()
    def forward(self, x_2d, x_3d, t):
        """x_2d (B,F,J,2), x_3d (B,H,F,J,3), t (B,) -> (B,H,F,J,3)  (mixste.py:278-298, eval branch)."""
        if self.is_train:
            raise NotImplementedError("training branch is outside the B200 inference path")
        if not x_3d.is_cuda:
            raise _native.PafuseError("pafuse_b200.MixSTE2 runs on CUDA tensors only (no CPU fallback)")
        tv = t.reshape(-1)
        t0 = float(tv[0].item())
        if tv.numel() > 1 and not bool((tv == tv[0]).all()):
            raise NotImplementedError("per-sample timesteps are a training-only feature of the reference")
        sinus = sinusoidal_embedding_cpu(t0, self.embed_dim).to(x_3d.device)
        return self._native(x_3d.device).pred_parts(x_2d, x_3d, sinus)
