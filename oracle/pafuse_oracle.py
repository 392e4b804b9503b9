"""CPU oracle for the PAFUSE denoising inference path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU fp32/fp64 tensor
arithmetic, no nn.Module, one fixed [S,F,J,C] activation layout) of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product path (``pafuse_b200``) never does and fails loudly without its CUDA
library.

Parity pin: the reference ships no tests for this path, so
the oracle is pinned against outputs of the reference itself, imported
unmodified from /root/reference by ``oracle/ref_harness.py`` and frozen as
``tests/golden/*.npz`` by ``tests/golden/make_golden.py`` (script + vectors are
committed).  ``tests/test_oracle_golden.py`` checks every function below against
those vectors; the one reference known-answer test (``common/utils.py:129-157``)
is restated in ``tests/test_oracle_golden.py::test_reference_known_answer_test_of_part_functions``.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as Fn

# --------------------------------------------------------------------------
# schedule (common/diffusionpose.py:41-51, 90-132, 279-281, 302-306)
# --------------------------------------------------------------------------


def cosine_alphas_cumprod(timesteps: int = 1000, s: float = 0.008) -> torch.Tensor:
    """fp64 alphas_cumprod, diffusionpose.py:41-51 and :92-93."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    return torch.cumprod(1.0 - betas, dim=0)


def sampling_times(total_timesteps: int, sampling_timesteps: int):
    """[(t, t_next)] pairs, diffusionpose.py:279-281."""
    times = torch.linspace(-1, total_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_coefficients(alphas_cumprod: torch.Tensor, t: int, t_next: int, eta: float = 1.0):
    """(sqrt_recip, sqrt_recipm1) fp64 and (sqrt(a_next), c, sigma) fp64 0-dim
    tensors for one step: diffusionpose.py:119-120, :302-306."""
    a = alphas_cumprod[t]
    sr = torch.sqrt(1.0 / a)
    srm1 = torch.sqrt(1.0 / a - 1)
    if t_next < 0:
        return sr, srm1, None, None, None
    an = alphas_cumprod[t_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return sr, srm1, an.sqrt(), c, sigma


# --------------------------------------------------------------------------
# denoiser (common/mixste.py)
# --------------------------------------------------------------------------


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """mixste.py:132-139."""
    half = dim // 2
    f = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    e = t[:, None] * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def time_embedding(w: dict, t: torch.Tensor, C: int) -> torch.Tensor:
    """time_mlp: sinusoid -> Linear(C,2C) -> GELU(erf) -> Linear(2C,C); mixste.py:179-184."""
    e = sinusoidal_embedding(t, C)
    e = Fn.gelu(Fn.linear(e, w["time_mlp.1.weight"], w["time_mlp.1.bias"]))
    return Fn.linear(e, w["time_mlp.3.weight"], w["time_mlp.3.bias"])


def _attention(w: dict, pre: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention.forward with comb=False, mixste.py:63-82.  x: (G, L, C)."""
    G, L, C = x.shape
    hd = C // heads
    qkv = Fn.linear(x, w[pre + "attn.qkv.weight"], w[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(G, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    a = a.softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(G, L, C)
    return Fn.linear(o, w[pre + "attn.proj.weight"], w[pre + "attn.proj.bias"])


def _block(w: dict, pre: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """Block.forward (changedim=False), mixste.py:113-116; norm eps 1e-6 (:163)."""
    C = x.shape[-1]
    x = x + _attention(w, pre, Fn.layer_norm(x, (C,), w[pre + "norm1.weight"], w[pre + "norm1.bias"], 1e-6), heads)
    h = Fn.layer_norm(x, (C,), w[pre + "norm2.weight"], w[pre + "norm2.bias"], 1e-6)
    h = Fn.gelu(Fn.linear(h, w[pre + "mlp.fc1.weight"], w[pre + "mlp.fc1.bias"]))
    return x + Fn.linear(h, w[pre + "mlp.fc2.weight"], w[pre + "mlp.fc2.bias"])


def mixste_forward(w: dict, x_2d: torch.Tensor, x_3d: torch.Tensor, t: torch.Tensor,
                   depth: int = 8, heads: int = 8) -> torch.Tensor:
    """MixSTE2.forward, inference branch (mixste.py:226-245, 247-258, 260-276, 278-298).

    x_2d (B,F,J,2), x_3d (B,H,F,J,3), t (B,) long -> (B,H,F,J,3).
    Activations stay in one [S=B*H, F, J, C] tensor; spatial blocks attend over
    J inside each (s,f), temporal blocks over F inside each (s,j).
    """
    B, H, F, J, _ = x_3d.shape
    C = w["Spatial_pos_embed"].shape[-1]
    feat = torch.cat((x_2d[:, None].expand(B, H, F, J, 2), x_3d), dim=-1)          # [u,v,x,y,z], :227-228
    x = Fn.linear(feat, w["Spatial_patch_to_embedding.weight"], w["Spatial_patch_to_embedding.bias"])
    x = x + w["Spatial_pos_embed"].reshape(1, 1, 1, J, C)
    x = x + time_embedding(w, t.float() if t.dtype != torch.float32 else t, C)[:, None, None, None, :]
    x = x.reshape(B * H, F, J, C)
    S = B * H
    sn = (w["Spatial_norm.weight"], w["Spatial_norm.bias"])
    tn = (w["Temporal_norm.weight"], w["Temporal_norm.bias"])
    for i in range(depth):
        # spatial block over joints, then the shared Spatial_norm (:239-243, :268-269)
        x = _block(w, f"STEblocks.{i}.", x.reshape(S * F, J, C), heads)
        x = Fn.layer_norm(x, (C,), sn[0], sn[1], 1e-6).reshape(S, F, J, C)
        # temporal block over frames, then the shared Temporal_norm (:249-257, :272-273)
        xt = x.permute(0, 2, 1, 3).reshape(S * J, F, C)
        if i == 0:
            xt = xt + w["Temporal_pos_embed"]
        xt = _block(w, f"TTEblocks.{i}.", xt, heads)
        xt = Fn.layer_norm(xt, (C,), tn[0], tn[1], 1e-6)
        x = xt.reshape(S, J, F, C).permute(0, 2, 1, 3)
    x = Fn.layer_norm(x, (C,), w["head.0.weight"], w["head.0.bias"], 1e-5)          # nn.LayerNorm default eps, :208
    x = Fn.linear(x, w["head.1.weight"], w["head.1.bias"])
    return x.reshape(B, H, F, J, 3)


def part_weights(state_dict: dict, part: str, prefix: str = "pose_estimator.") -> dict:
    p = f"{prefix}{part}."
    return {k[len(p):]: v for k, v in state_dict.items() if k.startswith(p)}


def pred_parts(state_dict: dict, parts: dict, x_2d, x_3d, t, depth=8, heads=8):
    """split per part -> denoise -> concatenate in part order; diffusionpose.py:163-172, :328-335."""
    outs = []
    for part, idx in parts.items():
        outs.append(mixste_forward(part_weights(state_dict, part), x_2d[..., idx, :], x_3d[..., idx, :], t, depth, heads))
    return torch.cat(outs, dim=-2)


# --------------------------------------------------------------------------
# sampler (common/diffusionpose.py:192-225, 272-316)
# --------------------------------------------------------------------------


def flip_pose(x: torch.Tensor, joints_left, joints_right) -> torch.Tensor:
    """negate x and swap left/right joints (joint axis = -2); diffusionpose.py:195-198, :211-213."""
    y = x.clone()
    y[..., 0] *= -1
    y[..., joints_left + joints_right, :] = y[..., joints_right + joints_left, :]
    return y


def model_predictions_flip(state_dict, parts, img, x2d, x2d_flip, t_int, joints_left, joints_right,
                           sr, srm1, scale=1.0, depth=8, heads=8):
    """x0 and eps for one step with flip-TTA; diffusionpose.py:192-225."""
    B = img.shape[0]
    t = torch.full((B,), t_int, dtype=torch.long)
    x_t = torch.clamp(img, min=-1.1 * scale, max=1.1 * scale) / scale
    x_t_flip = flip_pose(x_t, joints_left, joints_right)
    pred = pred_parts(state_dict, parts, x2d, x_t, t, depth, heads)
    pred_f = pred_parts(state_dict, parts, x2d_flip, x_t_flip, t, depth, heads)
    pred = (pred + flip_pose(pred_f, joints_left, joints_right)) / 2
    x0 = torch.clamp(pred * scale, min=-1.1 * scale, max=1.1 * scale)
    eps = ((sr * img.double() - x0.double()) / srm1).float()                      # fp64 then .float(), :157-161, :222-223
    return eps, x0


def ddim_update(x0, eps, noise, sqrt_an, c, sigma):
    """img = x0*sqrt(a_next) + c*eps + sigma*noise in fp32 (0-dim fp64 scalars do not promote); :310-312."""
    a32, c32, s32 = (torch.tensor(float(v), dtype=torch.float32) for v in (sqrt_an, c, sigma))
    return x0 * a32 + c32 * eps + s32 * noise


def ddim_sample_flip(state_dict, parts, x2d, x2d_flip, noises, joints_left, joints_right,
                     num_proposals, sampling_timesteps, total_timesteps=1000, scale=1.0, depth=8, heads=8,
                     return_trace=False):
    """DDIM loop with flip-TTA; diffusionpose.py:272-316.  ``noises`` = the tensors the
    reference would draw: noises[0] is the initial img (:283), noises[k] the k-th randn_like (:308)."""
    ac = cosine_alphas_cumprod(total_timesteps)
    img = noises[0]
    preds, trace, draw = [], [], 1
    for t, t_next in sampling_times(total_timesteps, sampling_timesteps):
        sr, srm1, san, c, sigma = ddim_coefficients(ac, t, t_next)
        eps, x0 = model_predictions_flip(state_dict, parts, img, x2d, x2d_flip, t, joints_left, joints_right,
                                         sr, srm1, scale, depth, heads)
        preds.append(x0)
        if t_next < 0:
            img = x0
            continue
        img = ddim_update(x0, eps, noises[draw], san, c, sigma)
        draw += 1
        trace.append(img)
    out = torch.stack(preds, dim=1)
    return (out, trace) if return_trace else out


def ddim_sample_noflip(state_dict, parts, x2d, noises, sampling_timesteps, total_timesteps=1000, scale=1.0,
                       depth=8, heads=8):
    """Non-TTA sampler, valid for num_proposals == 1 only (the reference raises for H>1,
    SURVEY.md 7.3); diffusionpose.py:174-190, :227-270."""
    ac = cosine_alphas_cumprod(total_timesteps)
    img = noises[0]
    assert img.shape[1] == 1
    B = img.shape[0]
    preds, draw = [], 1
    for t, t_next in sampling_times(total_timesteps, sampling_timesteps):
        sr, srm1, san, c, sigma = ddim_coefficients(ac, t, t_next)
        x_t = torch.clamp(img, min=-1.1 * scale, max=1.1 * scale) / scale
        pred = pred_parts(state_dict, parts, x2d, x_t, torch.full((B,), t, dtype=torch.long), depth, heads)
        x0 = torch.clamp(pred * scale, min=-1.1 * scale, max=1.1 * scale)
        eps = (sr * img.double() - x0.double()) / srm1                              # stays fp64 here (:189)
        preds.append(x0)
        if t_next < 0:
            img = x0
            continue
        img = (x0 * san + c * eps + sigma * noises[draw]).float()                   # :265-268
        draw += 1
    return torch.stack(preds, dim=1)


# --------------------------------------------------------------------------
# post-processing
# --------------------------------------------------------------------------


def wb_pose_from_parts(pose: torch.Tensor, parts_joint_indices: dict, connection: dict):
    """Part re-assembly, common/utils.py:113-126 with center_pose_at_root(revert=True) (:79-92).

    Returns (whole_body, input_after_call): the reference negates the root rows
    0/1/10/11 of its *input* in place (offset is a view), which makes
    out[root] = (-x) + x = +0.0 and out[j] = x[j] + x[root] otherwise.
    """
    conn = dict(connection)
    conn["body"] = 0
    x = pose.clone()
    out = torch.zeros_like(pose)
    for part, idx in parts_joint_indices.items():
        if part in conn:
            r = conn[part]
            x[..., r, :] = -x[..., r, :]
            out[..., idx, :] = (x - (x[..., r:r + 1, :]))[..., idx, :]
    return out, x


def center_pose_parts(pose: torch.Tensor, parts_joint_indices: dict, root_indices: dict) -> torch.Tensor:
    """Part-centred pose (applied to the GROUND TRUTH by the callers, main_h3wb.py:304):
    out[part joints] = x[part joints] - x[root(part)]; common/utils.py:95-110 with
    center_pose_at_root (:79-92, revert=False: no aliasing side effect)."""
    out = torch.zeros_like(pose)
    for part, idx in parts_joint_indices.items():
        r = root_indices[part]
        out[..., idx, :] = (pose - pose[..., r:r + 1, :])[..., idx, :]
    return out


def project_to_2d(X: torch.Tensor, cam: torch.Tensor) -> torch.Tensor:
    """H36M projection with distortion, common/camera.py:30-60.  X (N,*,3), cam (N,9)."""
    while cam.dim() < X.dim():
        cam = cam.unsqueeze(1)
    f, c, k, p = cam[..., :2], cam[..., 2:4], cam[..., 4:7], cam[..., 7:]
    XX = torch.clamp(X[..., :2] / X[..., 2:], min=-1, max=1)
    r2 = torch.sum(XX ** 2, dim=-1, keepdim=True)
    radial = 1 + torch.sum(k * torch.cat((r2, r2 ** 2, r2 ** 3), dim=-1), dim=-1, keepdim=True)
    tan = torch.sum(p * XX, dim=-1, keepdim=True)
    return f * (XX * (radial + tan) + p * r2) + c


def aggregate(pred: torch.Tensor, traj: torch.Tensor, cam: torch.Tensor, x2d: torch.Tensor):
    """J-Agg pose (per-joint argmin of 2D reprojection error over hypotheses,
    common/loss.py:90-108 + common/visualization.py:453-463) and P-Agg pose
    (mean over hypotheses, loss.py:68-70), with the reprojection of
    main_h3wb.py:336-342.

    pred (B,K,H,F,J,3) whole-body root-relative, traj (B,F,1,3), cam (1,9), x2d (B,F,J,2)
    -> jagg (B,K,F,J,3), pagg (B,K,F,J,3), select (B,K,F,J) int64.
    """
    B, K, H, F, J, _ = pred.shape
    absd = pred + traj[:, None, None]
    reproj = project_to_2d(absd.reshape(B * K * H * F, J, 3), cam.repeat(B * K * H * F, 1)).reshape(B, K, H, F, J, 2)
    err = torch.norm(reproj - x2d[:, None, None], dim=-1)                          # (B,K,H,F,J)
    sel = torch.min(err, dim=2, keepdim=True).indices                              # first minimum
    jagg = torch.gather(pred, 2, sel.unsqueeze(-1).expand(B, K, 1, F, J, 3)).squeeze(2)
    pagg = torch.mean(pred, dim=2)
    return jagg, pagg, sel.squeeze(2)


# --------------------------------------------------------------------------
# caller-side prep (main_h3wb.py:122-154, 268-270; in_the_wild/h3wb_diffusion.py:119-133)
# --------------------------------------------------------------------------


def eval_data_prepare(receptive_field: int, seq: torch.Tensor) -> torch.Tensor:
    """(T,J,C) -> (ceil(T/rf), rf, J, C); last clip right-aligned, T<rf replicate-padded."""
    T = seq.shape[0]
    n = (T + receptive_field - 1) // receptive_field
    out = torch.empty((n, receptive_field) + tuple(seq.shape[1:]), dtype=seq.dtype)
    for i in range(n - 1):
        out[i] = seq[i * receptive_field:(i + 1) * receptive_field]
    if T < receptive_field:
        seq = torch.cat((seq, seq[-1:].expand((receptive_field - T,) + tuple(seq.shape[1:]))), dim=0)
    out[-1] = seq[-receptive_field:]
    return out


def stitch_clips(pred: torch.Tensor, total_frames: int) -> torch.Tensor:
    """(N,K,H,rf,J,3) -> (K,H,T,J,3); h3wb_diffusion.py:119-133."""
    N, K, H, rf, J, _ = pred.shape
    out = torch.empty((K, H, total_frames, J, 3), dtype=pred.dtype)
    full = total_frames // rf
    for i in range(full):
        out[:, :, i * rf:(i + 1) * rf] = pred[i]
    left = total_frames - full * rf
    if left > 0:
        out[:, :, -left:] = pred[-1][:, :, -left:]
    return out


def flip_inputs_2d(x2d: torch.Tensor, kps_left, kps_right) -> torch.Tensor:
    """Flip-TTA twin of a 2D input built by the callers: main_h3wb.py:268-270, in_the_wild/utils.py:340-342."""
    out = x2d.clone()
    out[..., 0] *= -1
    out[..., list(kps_left) + list(kps_right), :] = out[..., list(kps_right) + list(kps_left), :]
    return out


def normalize_screen_coordinates(X, w: int, h: int):
    """common/camera.py:7-11 on a numpy float32 array: X / w * 2 stays float32, the subtraction of the python-float
    list promotes to float64 (that is what the reference hands to astype('float32') later)."""
    import numpy as np
    assert X.shape[-1] == 2
    return X / w * 2 - np.array([1, h / w])


def keypoints_from_openpifpaf(detections, w: int, h: int) -> torch.Tensor:
    """in_the_wild/h3wb_diffusion.py:64-77 (+ the astype('float32') of in_the_wild/utils.py:338): detections
    (T,133,3) pixel (x, y, confidence) -> (T,134,2) normalised input with joint 0 = mean of joints 12 and 13."""
    import numpy as np
    det = detections.numpy().astype(np.float32)
    T = det.shape[0]
    kp = np.zeros((T, det.shape[1] + 1, 2), dtype=np.float32)
    kp[:, 1:, 0] = det[:, :, 0]
    kp[:, 1:, 1] = det[:, :, 1]
    kp[:, :1, :] = (kp[:, 12:13, :] + kp[:, 13:14, :]) / 2.
    return torch.from_numpy(normalize_screen_coordinates(kp, w, h).astype("float32"))


# --------------------------------------------------------------------------
# GT-dependent multi-hypothesis metrics (common/loss.py:36-146, whole-body calls of main_h3wb.py:344-349)
# --------------------------------------------------------------------------
def mpjpe_j_best(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mpjpe_diffusion_all_min(mean_pos=False), loss.py:53-66: (B,K,H,F,J,3), (B,F,J,3) -> (K,)."""
    err = torch.norm(pred - target[:, None, None], dim=-1)            # b k h f j
    return err.min(dim=2).values.permute(1, 0, 2, 3).reshape(pred.shape[1], -1).mean(dim=-1)


def mpjpe_p_agg(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mpjpe_diffusion_all_min(mean_pos=True), loss.py:68-76."""
    err = torch.norm(pred.mean(dim=2) - target[:, None], dim=-1)      # b k f j
    return err.permute(1, 0, 2, 3).reshape(pred.shape[1], -1).mean(dim=-1)


def mpjpe_j_agg(pred: torch.Tensor, target: torch.Tensor, reproj_2d: torch.Tensor, target_2d: torch.Tensor) -> torch.Tensor:
    """mpjpe_diffusion_reproj, loss.py:90-112."""
    err = torch.norm(pred - target[:, None, None], dim=-1)
    err2d = torch.norm(reproj_2d - target_2d[:, None, None], dim=-1)
    sel = err2d.min(dim=2, keepdim=True).indices
    picked = torch.gather(err, 2, sel)
    return picked.permute(1, 2, 0, 3, 4).reshape(pred.shape[1], -1).mean(dim=-1)


def mpjpe_p_best(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mpjpe_diffusion(mean_pos=False, part_based=False), loss.py:114-146 (both poses root-centred first, :131-132)."""
    p = pred - pred[..., 0:1, :]
    t = target - target[..., 0:1, :]
    err = torch.norm(p - t[:, None, None], dim=-1)                    # b k h f j
    K, H = pred.shape[1], pred.shape[2]
    return err.permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(dim=-1).min(dim=1).values


def mpjpe_p_best_parts(pred: torch.Tensor, target: torch.Tensor, parts_joint_indices: dict, root_indices: dict):
    """mpjpe_diffusion(mean_pos=False, part_based=True), loss.py:119-127,135-154: both poses centred per part, the
    hypothesis with the smallest whole-pose mean error wins, its per-part means are returned next to that minimum."""
    p = center_pose_parts(pred, parts_joint_indices, root_indices)
    t = center_pose_parts(target, parts_joint_indices, root_indices)
    err = torch.norm(p - t[:, None, None], dim=-1)                    # b k h f j
    K, H = pred.shape[1], pred.shape[2]
    per_h = err.permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(dim=-1)
    best, inds = per_h.min(dim=1)
    parts = {}
    for name, idx in parts_joint_indices.items():
        e = err[..., idx].permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(dim=-1)
        parts[name] = e.gather(1, inds.view(-1, 1)).squeeze(1)
    return best, parts


def mpjpe_p_agg_parts(pred: torch.Tensor, target: torch.Tensor, parts_joint_indices: dict, root_indices: dict):
    """mpjpe_diffusion_all_min(mean_pos=True, part_based=True), loss.py:41-51,68-86."""
    p = center_pose_parts(pred, parts_joint_indices, root_indices)
    t = center_pose_parts(target, parts_joint_indices, root_indices)
    err = torch.norm(p.mean(dim=2) - t[:, None], dim=-1)              # b k f j
    K = pred.shape[1]
    parts = {name: err[..., idx].permute(1, 0, 2, 3).reshape(K, -1).mean(dim=-1) for name, idx in parts_joint_indices.items()}
    return err.permute(1, 0, 2, 3).reshape(K, -1).mean(dim=-1), parts
