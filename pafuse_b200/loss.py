"""GT-dependent multi-hypothesis MPJPE protocols of the reference's ``evaluate()`` on the device.

Same names and argument meaning as ``common/loss.py`` for the whole-body (non part-based) calls that
``main_h3wb.py:344-349`` makes; all of them come out of one kernel pass (``pafuse_mpjpe_metrics``):

    mpjpe_diffusion_all_min(pred, target)                 J-Best   loss.py:53-66
    mpjpe_diffusion_all_min(pred, target, mean_pos=True)  P-Agg    loss.py:68-76
    mpjpe_diffusion_reproj(pred, target, reproj, x2d)     J-Agg    loss.py:90-112
    mpjpe_diffusion(pred, target)                         P-Best   loss.py:114-146

``pred`` (B,K,H,F,J,3), ``target`` (B,F,J,3); results are (K,) tensors (fp32 like the reference; the device sums are
fp64).  The part-based variants (``part_based=True``) and the Procrustes metrics are not implemented.
"""
from __future__ import annotations

import torch

from . import _native
from .utils import _post_context

__all__ = ["evaluate_metrics", "mpjpe_diffusion_all_min", "mpjpe_diffusion_reproj", "mpjpe_diffusion"]


def _means(pred, target, traj=None, cam=None, x2d=None, reproj=None):
    if not pred.is_cuda:
        raise _native.PafuseError("pafuse_b200.loss needs CUDA tensors (no CPU fallback)")
    B, K, H, F, J, _ = pred.shape
    ctx = _post_context(pred.device, J, F)
    if x2d is None:                                            # protocols that do not look at the 2D error
        x2d = torch.zeros((B, F, J, 2), dtype=torch.float32, device=pred.device)
    if reproj is None and cam is None:
        reproj = torch.zeros((B, K, H, F, J, 2), dtype=torch.float32, device=pred.device)
    return ctx.mpjpe_metrics(pred, target, traj, cam, x2d, reproj)      # (K, 3 + H) means


def evaluate_metrics(pred, target, inputs_traj, cam, inputs_2d):
    """All four protocols of ``main_h3wb.py:336-349`` (reprojection of ``pred + traj`` included): dict of (K,) tensors."""
    m = _means(pred, target, inputs_traj, cam, inputs_2d)
    return {"J-Best": m[:, 0].float(), "P-Agg": m[:, 1].float(), "J-Agg": m[:, 2].float(),
            "P-Best": m[:, 3:].min(dim=1).values.float()}


def mpjpe_diffusion_all_min(predicted, target, mean_pos=False, part_based=False, dataset=None):
    if part_based:
        raise NotImplementedError("part-based MPJPE variants are not implemented on the device")
    m = _means(predicted, target)
    return m[:, 1].float() if mean_pos else m[:, 0].float()


def mpjpe_diffusion_reproj(predicted, target, reproj_2d, target_2d):
    return _means(predicted, target, x2d=target_2d, reproj=reproj_2d)[:, 2].float()


def mpjpe_diffusion(predicted, target, mean_pos=False, part_based=False, dataset=None):
    if part_based:
        raise NotImplementedError("part-based MPJPE variants are not implemented on the device")
    m = _means(predicted, target)
    if mean_pos:
        raise NotImplementedError("mpjpe_diffusion(mean_pos=True) is not used by evaluate(); use mpjpe_diffusion_all_min")
    return m[:, 3:].min(dim=1).values.float(), {}
