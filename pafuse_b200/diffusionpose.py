"""Drop-in replacement for the reference ``D3DP`` module (inference path).

Keeps the reference's constructor, buffers, ``state_dict`` keys and
``forward(input_2d, input_3d, input_2d_flip=None)`` contract
(``common/diffusionpose.py:54-153, 337-344``); the DDIM loop
(``ddim_sample_flip`` :272-316 / ``ddim_sample`` :227-270) drives the sm_100a
library through the C ABI, one ``pafuse_ddim_step`` call per sampling step.

Differences that are deliberate:
  * training (``is_train=True`` forward) is out of scope and raises;
  * the non-TTA sampler also works for ``num_proposals > 1`` (the reference
    raises an EinopsError there);
  * ``noise_source`` lets a caller inject the Gaussian draws (parity tests,
    multi-GPU sharding of one global draw); by default the draws are
    ``torch.randn`` on the input's device in the reference's order and shapes.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _native
from .h3wb import flip_permutation
from .mixste import MixSTE2, sinusoidal_embedding_cpu, state_items, weights_fingerprint

__all__ = ["D3DP"]

PART_CHANNELS = {"body": 384, "face": 224, "hands": 256}   # diffusionpose.py:141


def cosine_beta_schedule(timesteps, s=0.008):
    """Cosine schedule in fp64, same operations as diffusionpose.py:41-51 so the
    registered buffers are bit-identical to a reference checkpoint's."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


class D3DP(nn.Module):
    def __init__(self, args, joints_left, joints_right, dataset, is_train=True, num_proposals=1, sampling_timesteps=1):
        super().__init__()
        self.args = args
        self.frames = args.model.number_of_frames
        self.num_proposals = num_proposals
        self.flip = args.model.test_time_augmentation
        self.joints_left = list(joints_left)
        self.joints_right = list(joints_right)
        self.is_train = is_train
        self.num_kps = args.data.num_kps
        self.diff_model = args.model.diff_model
        self.device = "cuda"
        self.dataset = dataset
        self.metadata = dataset.metadata
        self.parts_root_indices = dataset.root_indices
        pji = dataset.parts_joint_indices.copy()
        if args.data.merge_hands:                      # diffusionpose.py:76-83
            pji["hands"] = pji["left_hand"] + pji["right_hand"]
            del pji["left_hand"], pji["right_hand"]
        self.parts_joint_indices = pji

        if self.diff_model != "MixSTE2":
            raise Exception(f"The model {self.diff_model} does not exist")
        if not args.general.part_based_model:
            raise NotImplementedError("pafuse_b200 implements the part-based PAFUSE model (general.part_based_model=True)")

        timesteps = args.ft2d.timestep
        betas = cosine_beta_schedule(timesteps)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = 1.
        self.self_condition = False
        self.scale = args.ft2d.scale
        self.objective = "pred_x0"

        # the 12 fp64 schedule buffers of the reference state_dict (diffusionpose.py:107-132)
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        for name, val in (
            ("betas", betas),
            ("alphas_cumprod", alphas_cumprod),
            ("alphas_cumprod_prev", alphas_cumprod_prev),
            ("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod)),
            ("sqrt_one_minus_alphas_cumprod", torch.sqrt(1. - alphas_cumprod)),
            ("log_one_minus_alphas_cumprod", torch.log(1. - alphas_cumprod)),
            ("sqrt_recip_alphas_cumprod", torch.sqrt(1. / alphas_cumprod)),
            ("sqrt_recipm1_alphas_cumprod", torch.sqrt(1. / alphas_cumprod - 1)),
            ("posterior_variance", posterior_variance),
            ("posterior_log_variance_clipped", torch.log(posterior_variance.clamp(min=1e-20))),
            ("posterior_mean_coef1", betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod)),
            ("posterior_mean_coef2", (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod)),
        ):
            self.register_buffer(name, val)

        drop_path_rate = 0.1 if is_train else 0
        self.pose_estimator = nn.ModuleDict({
            part: MixSTE2(num_frame=self.frames, num_joints=len(idx), in_chans=args.model.input_size,
                          embed_dim_ratio=PART_CHANNELS[part], depth=args.model.dep, num_heads=8, mlp_ratio=2.,
                          qkv_bias=True, qk_scale=None, drop_path_rate=drop_path_rate, is_train=is_train)
            for part, idx in self.parts_joint_indices.items()
        })

        self.noise_source = None          # optional callable(k, shape, device) -> tensor
        self.max_seqs = int(getattr(getattr(args, "b200", None), "max_seqs", 0) or 640)
        self._natives = {}                # device index -> (context, fingerprint of the weights it holds)
        self._native_dirty = False        # set to force a rebuild (e.g. after changing max_seqs)
        self._sinus_cache = {}
        self._schedule_cpu = None

    # ------------------------------------------------------------------ plumbing
    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts reference checkpoints, including DataParallel's ``module.`` prefix
        (``main_h3wb.py:711-714``: keys are ``module.pose_estimator...``)."""
        if any(k.startswith("module.") for k in state_dict):
            state_dict = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._schedule_cpu = None
        return super()._apply(fn, *a, **k)

    def _replicate_for_data_parallel(self):
        """``nn.DataParallel`` (the reference's multi-GPU mode, main_h3wb.py:699-705) re-broadcasts the weights into
        fresh tensors for every forward; the replica carries the SOURCE module's fingerprint so that its device's
        context is rebuilt when the source weights change, not on every forward."""
        replica = super()._replicate_for_data_parallel()
        replica._source_fp = self._fingerprint()
        return replica

    def _fingerprint(self):
        fp = getattr(self, "_source_fp", None)
        return fp if fp is not None else weights_fingerprint(self.pose_estimator)

    def _native(self, device) -> _native.NativeContext:
        key = torch.device(device).index
        if key is None:
            key = torch.cuda.current_device()
        if self._native_dirty:
            for ctx, _ in self._natives.values():
                ctx.close()
            self._natives.clear()
            self._native_dirty = False
        # the packed fp16 hi/lo copies inside the library follow ANY change of the weights: load_state_dict on the
        # module or on a child, in-place edits, .to() / .cuda() (new storage)
        fp = self._fingerprint()
        held = self._natives.get(key)
        if held is not None and held[1] != fp:
            held[0].close()
            del self._natives[key]
        if key not in self._natives:
            parts = list(self.parts_joint_indices.items())
            ctx = _native.NativeContext(
                self.frames, self.num_kps, self.args.model.dep, 8,
                [PART_CHANNELS[p] for p, _ in parts], [list(idx) for _, idx in parts],
                flip_permutation(self.joints_left, self.joints_right, self.num_kps),
                self.scale, self.max_seqs, torch.device("cuda", key))
            for pi, (part, _) in enumerate(parts):
                for name, t in state_items(self.pose_estimator[part]):
                    ctx.set_weight(pi, name, t)
            ctx.commit_weights()
            self._natives[key] = (ctx, fp)
        return self._natives[key][0]

    def native_context(self, device=None):
        """The per-device C-ABI context (used by the post-processing helpers and bench)."""
        return self._native(device if device is not None else torch.device("cuda", torch.cuda.current_device()))

    def _sinus(self, t: int, device):
        key = (int(t), str(device))
        if key not in self._sinus_cache:
            tab = torch.cat([sinusoidal_embedding_cpu(t, PART_CHANNELS[p]) for p in self.parts_joint_indices])
            self._sinus_cache[key] = tab.to(device)
        return self._sinus_cache[key]

    def _schedule(self):
        if self._schedule_cpu is None:
            self._schedule_cpu = (self.alphas_cumprod.detach().double().cpu(),
                                  self.sqrt_recip_alphas_cumprod.detach().double().cpu(),
                                  self.sqrt_recipm1_alphas_cumprod.detach().double().cpu())
        return self._schedule_cpu

    def sampling_time_pairs(self):
        """[(t, t_next)], diffusionpose.py:279-281."""
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def step_coefficients(self, t, t_next):
        """fp64 (sqrt_recip, sqrt_recipm1) at t and fp64 (sqrt(a_next), c, sigma) of diffusionpose.py:302-306."""
        ac, sr, srm1 = self._schedule()
        if t_next < 0:
            return float(sr[t]), float(srm1[t]), 0.0, 0.0, 0.0
        alpha, alpha_next = ac[t], ac[t_next]
        sigma = self.ddim_sampling_eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        return float(sr[t]), float(srm1[t]), float(alpha_next.sqrt()), float(c), float(sigma)

    def _draw(self, k, shape, device):
        if self.noise_source is not None:
            n = self.noise_source(k, shape, device)
            return n.to(device=device, dtype=torch.float32).contiguous()
        return torch.randn(shape, device=device)

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def ddim_sample_flip(self, inputs_2d, inputs_3d=None, clip_denoised=True, do_postprocess=True, input_2d_flip=None,
                         wb_preds=True):
        return self._sample(inputs_2d, input_2d_flip, flip=True)

    @torch.no_grad()
    def ddim_sample(self, inputs_2d, inputs_3d=None, clip_denoised=True, do_postprocess=True, wb_preds=True):
        return self._sample(inputs_2d, None, flip=False)

    def _sample(self, inputs_2d, input_2d_flip, flip):
        if not inputs_2d.is_cuda:
            raise _native.PafuseError("pafuse_b200.D3DP runs on CUDA tensors only (no CPU fallback); "
                                      "move the module and its inputs to a B200 device")
        if flip and input_2d_flip is None:
            raise ValueError("test_time_augmentation is on: forward() needs input_2d_flip")
        device = inputs_2d.device
        ctx = self._native(device)
        B = inputs_2d.shape[0]
        H, K, Fr, J = self.num_proposals, self.sampling_timesteps, self.frames, self.num_kps
        assert tuple(inputs_2d.shape[1:]) == (Fr, J, 2), f"input_2d must be (B,{Fr},{J},2), got {tuple(inputs_2d.shape)}"
        x2d = inputs_2d.detach().to(torch.float32).contiguous()
        x2d_flip = input_2d_flip.detach().to(device=device, dtype=torch.float32).contiguous() if flip else None
        shape = (B, H, Fr, J, 3)
        if B == 0:
            return torch.empty((0, K, H, Fr, J, 3), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            img = self._draw(0, shape, device)                        # diffusionpose.py:283
            if self.noise_source is not None:
                img = img.clone()                                      # the state is updated in place
            preds_all = torch.empty((B, K, H, Fr, J, 3), dtype=torch.float32, device=device)
            stride_b = K * H * Fr * J * 3
            step_elems = H * Fr * J * 3
            for k, (t, t_next) in enumerate(self.sampling_time_pairs()):
                last = t_next < 0
                sr, srm1, san, c, sigma = self.step_coefficients(t, t_next)
                noise = None if last else self._draw(k + 1, shape, device)     # :308 (drawn after the model call there;
                # the values do not depend on the order because nothing else consumes the generator in between)
                ctx.ddim_step(x2d, x2d_flip, self._sinus(t, device), img, noise,
                              preds_all.data_ptr() + 4 * k * step_elems, stride_b, B, H, flip, last,
                              sr, srm1, c, san, c, sigma)
        return preds_all

    def forward(self, input_2d, input_3d, input_2d_flip=None):
        if self.is_train:
            raise NotImplementedError("pafuse_b200.D3DP implements the inference path only (is_train=False)")
        if self.flip:
            return self.ddim_sample_flip(input_2d, input_3d, input_2d_flip=input_2d_flip)
        return self.ddim_sample(input_2d, input_3d)

    # sub-boundary kept for callers of the reference method (diffusionpose.py:163-172)
    @torch.no_grad()
    def pred_parts(self, inputs_2d, inputs_3d, t):
        tv = t.reshape(-1)
        return self._native(inputs_3d.device).pred_parts(inputs_2d, inputs_3d, self._sinus(int(tv[0].item()), inputs_3d.device))
