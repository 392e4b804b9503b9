"""Deterministic synthetic configs, weights and inputs (no dataset / checkpoint).

``pafuse_model.bin`` and the H3WB npz files are release assets that are not
available offline (reference ``README.md:43``, ``common/h3wb_dataset.py:18-23``),
so tests and ``bench.py`` use random-init weights and synthetic 2D input of the
reference's shapes.  Everything is generated on the CPU with explicitly seeded
``torch.Generator`` objects, keyed by tensor *name*, so that the reference, the
oracle and the CUDA path can be handed identical values on any machine.
"""
from __future__ import annotations

import hashlib
import math
from types import SimpleNamespace

import torch

from .h3wb import H3WBSkeleton, merged_part_indices

PART_CHANNELS = {"body": 384, "face": 224, "hands": 256}  # diffusionpose.py:141


def default_args(number_of_frames=27, depth=8, test_time_augmentation=True, scale=1.0,
                 timestep=1000, num_kps=134, batch_size=1024):
    """Attribute-style config with the keys ``D3DP.__init__`` reads
    (``config/config.yaml`` names; ``diffusionpose.py:62-103,140-153``)."""
    return SimpleNamespace(
        general=SimpleNamespace(part_based_model=True),
        data=SimpleNamespace(num_kps=num_kps, merge_hands=True),
        model=SimpleNamespace(number_of_frames=number_of_frames, test_time_augmentation=test_time_augmentation,
                              diff_model="MixSTE2", input_size=5, dep=depth, cs=288, batch_size=batch_size),
        ft2d=SimpleNamespace(timestep=timestep, scale=scale, sampling_timesteps=5, num_proposals=10),
    )


def _gen_for(name: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def part_state_shapes(C: int, J: int, F: int, depth: int, in_chans: int = 5):
    """name -> shape for one MixSTE2 denoiser (``common/mixste.py:141-210``)."""
    shapes = {
        "Spatial_patch_to_embedding.weight": (C, in_chans),
        "Spatial_patch_to_embedding.bias": (C,),
        "Spatial_pos_embed": (1, J, C),
        "Temporal_pos_embed": (1, F, C),
        "time_mlp.1.weight": (2 * C, C), "time_mlp.1.bias": (2 * C,),
        "time_mlp.3.weight": (C, 2 * C), "time_mlp.3.bias": (C,),
        "Spatial_norm.weight": (C,), "Spatial_norm.bias": (C,),
        "Temporal_norm.weight": (C,), "Temporal_norm.bias": (C,),
        "head.0.weight": (C,), "head.0.bias": (C,),
        "head.1.weight": (3, C), "head.1.bias": (3,),
    }
    for stack in ("STEblocks", "TTEblocks"):
        for i in range(depth):
            p = f"{stack}.{i}."
            shapes[p + "norm1.weight"] = (C,)
            shapes[p + "norm1.bias"] = (C,)
            shapes[p + "attn.qkv.weight"] = (3 * C, C)
            shapes[p + "attn.qkv.bias"] = (3 * C,)
            shapes[p + "attn.proj.weight"] = (C, C)
            shapes[p + "attn.proj.bias"] = (C,)
            shapes[p + "norm2.weight"] = (C,)
            shapes[p + "norm2.bias"] = (C,)
            shapes[p + "mlp.fc1.weight"] = (2 * C, C)
            shapes[p + "mlp.fc1.bias"] = (2 * C,)
            shapes[p + "mlp.fc2.weight"] = (C, 2 * C)
            shapes[p + "mlp.fc2.bias"] = (C,)
    return shapes


def synthetic_state_dict(seed: int = 1, depth: int = 8, frames: int = 27, skeleton=None, prefix="pose_estimator."):
    """Random-init weights for the three part denoisers, keyed like the
    reference ``state_dict`` (``pose_estimator.<part>.<name>``).

    Linear weights/biases follow ``nn.Linear``'s default U(-1/sqrt(in), 1/sqrt(in));
    norm scales are 1 + 0.1 N(0,1) and shifts 0.1 N(0,1); positional embeddings
    (zero-initialised in the reference, ``mixste.py:171,174``) are drawn from
    N(0, 0.02^2) so that a joint/frame permutation bug cannot hide.
    """
    skeleton = skeleton or H3WBSkeleton()
    parts = merged_part_indices(skeleton.parts_joint_indices)
    sd = {}
    for part, idx in parts.items():
        C = PART_CHANNELS[part]
        shapes = part_state_shapes(C, len(idx), frames, depth)
        for name, shape in shapes.items():
            g = _gen_for(f"{part}.{name}", seed)
            if name.endswith("pos_embed"):
                t = torch.randn(shape, generator=g) * 0.02
            elif "norm" in name or name.startswith("head.0"):
                t = torch.randn(shape, generator=g) * 0.1
                if name.endswith("weight"):
                    t = t + 1.0
            else:
                fan_in = shapes[name.replace(".bias", ".weight")][1]
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            sd[f"{prefix}{part}.{name}"] = t.float().contiguous()
    return sd


def synthetic_inputs(B: int, seed: int = 1, frames: int = 27, skeleton=None):
    """``inputs_2d`` ~ U(-1,1) (normalised screen coords, ``camera.py:7-11``) and its
    flip-TTA twin built exactly like ``main_h3wb.py:268-270``."""
    skeleton = skeleton or H3WBSkeleton()
    g = _gen_for("inputs_2d", seed)
    x2d = torch.rand((B, frames, skeleton.num_kps, 2), generator=g) * 2 - 1
    return x2d.float(), flip_inputs_2d(x2d.float(), skeleton.joints_left, skeleton.joints_right)


def flip_inputs_2d(x2d, kps_left, kps_right):
    out = x2d.clone()
    out[..., 0] *= -1
    out[..., kps_left + kps_right, :] = out[..., kps_right + kps_left, :]
    return out


def synthetic_noise(B: int, H: int, K: int, seed: int = 1, frames: int = 27, num_kps: int = 134):
    """The K tensors the sampler draws (initial ``img`` + one per non-final step,
    ``diffusionpose.py:283,308``) as a list of ``(B,H,F,134,3)`` fp32 tensors."""
    g = _gen_for("noise", seed)
    return [torch.randn((B, H, frames, num_kps, 3), generator=g).float() for _ in range(K)]


# H36M camera 0 intrinsics normalised as in h3wb_dataset.py:104-114 (values: h36m_dataset.py:20-29)
def h36m_cam0_intrinsics():
    w, h = 1000.0, 1002.0
    cx, cy = 512.54150390625, 515.4514770507812
    fx, fy = 1145.0494384765625, 1143.7811279296875
    return torch.tensor([[fx / w * 2, fy / w * 2, cx / w * 2 - 1, cy / w * 2 - h / w,
                          -0.20709891617298126, 0.24777518212795258, -0.0030751503072679043,
                          -0.0009756988729350269, -0.00142447161488235]], dtype=torch.float32)


def synthetic_trajectory(B: int, seed: int = 1, frames: int = 27):
    """Root trajectory (B,F,1,3) with depth in [3,6] m so reprojection is well conditioned."""
    g = _gen_for("traj", seed)
    t = torch.rand((B, frames, 1, 3), generator=g)
    t[..., :2] = t[..., :2] * 1.0 - 0.5
    t[..., 2] = t[..., 2] * 3.0 + 3.0
    return t.float()
