"""Free functions of the reference's post-processing, same names and argument
meaning (``common/utils.py:113-126``, ``common/camera.py:30-60``) plus the
multi-hypothesis aggregation the callers apply (``common/loss.py:68-70,101-108``,
``common/visualization.py:453-463``, reprojection ``main_h3wb.py:336-342``),
each one a single sm_100a kernel behind the C ABI.
"""
from __future__ import annotations

import torch

from . import _native
from .h3wb import H3WBSkeleton

_contexts = {}


def _post_context(device, num_kps=134, frames=27, flip_perm=None) -> _native.NativeContext:
    """A weight-less context for the pre-/post-processing kernels on ``device`` (``flip_perm``: the left/right
    joint permutation the flip-TTA input construction uses; identity when not given)."""
    dev = torch.device(device)
    perm = tuple(flip_perm) if flip_perm is not None else tuple(range(num_kps))
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), num_kps, frames, perm)
    if key not in _contexts:
        _contexts[key] = _native.NativeContext(frames, num_kps, 1, 8, [32], [[0]], list(perm), 1.0, 1,
                                               torch.device("cuda", key[0]))
    return _contexts[key]


def connection_table(dataset, num_kps):
    """conn[j] = body joint the part of j hangs from (0 for the body itself); -1 if j is in no part."""
    conn_idx = dict(dataset.parts_connection_indices)
    conn_idx["body"] = 0                                   # utils.py:116
    conn = [-1] * num_kps
    for part, idx in dataset.parts_joint_indices.items():
        if part in conn_idx:
            for j in idx:
                conn[j] = conn_idx[part]
    return conn


def wb_pose_from_parts(part_based_pose, dataset, mutate_input=True):
    """Whole-body pose from part-centred predictions (utils.py:113-126).

    Like the reference, the call also (a) adds ``'body': 0`` to
    ``dataset.parts_connection_indices`` and (b) negates the connection rows
    0/1/10/11 of ``part_based_pose`` in place (``center_pose_at_root`` negates a
    view of its input); pass ``mutate_input=False`` to skip (b).
    """
    if not part_based_pose.is_cuda:
        raise _native.PafuseError("pafuse_b200.wb_pose_from_parts needs a CUDA tensor (no CPU fallback)")
    num_kps = part_based_pose.shape[-2]
    conn = connection_table(dataset, num_kps)
    dataset.parts_connection_indices.update({"body": 0})
    x = part_based_pose
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.to(torch.float32).contiguous()
        mutate_input = False                               # would not be visible to the caller anyway
    return _post_context(x.device, num_kps).wb_pose_from_parts(x, conn, mutate_input)


def project_to_2d(X, camera_params):
    """H36M projection with distortion (camera.py:30-60).  X (N,*,3), camera_params (N,9)."""
    assert X.shape[-1] == 3
    assert len(camera_params.shape) == 2
    assert camera_params.shape[-1] == 9
    assert X.shape[0] == camera_params.shape[0]
    if not X.is_cuda:
        raise _native.PafuseError("pafuse_b200.project_to_2d needs a CUDA tensor (no CPU fallback)")
    return _post_context(X.device).project_to_2d(X, camera_params)


def aggregate_hypotheses(pred, inputs_traj, cam, inputs_2d, return_select=False, return_reproj=False):
    """J-Agg and P-Agg poses of whole-body predictions.

    pred (B,K,H,F,J,3), inputs_traj (B,F,1,3) or None, cam (1,9) or (B,9), inputs_2d (B,F,J,2)
    -> (jagg, pagg[, select][, reproj]) with jagg/pagg (B,K,F,J,3).
    """
    if not pred.is_cuda:
        raise _native.PafuseError("pafuse_b200.aggregate_hypotheses needs CUDA tensors (no CPU fallback)")
    ctx = _post_context(pred.device, pred.shape[-2], pred.shape[-3])
    jagg, pagg, sel, rep = ctx.aggregate(pred, inputs_traj, cam, inputs_2d, want_select=return_select,
                                         want_reproj=return_reproj)
    out = [jagg, pagg]
    if return_select:
        out.append(sel)
    if return_reproj:
        out.append(rep)
    return tuple(out)


def eval_data_prepare(receptive_field, inputs_2d, kps_left=None, kps_right=None):
    """Tile one 2D sequence into clips (``main_h3wb.py:122-154``, ``in_the_wild/utils.py:279-320``).

    inputs_2d ``(1,T,J,2)`` or ``(T,J,2)`` CUDA tensor -> ``(ceil(T/rf), rf, J, 2)``: the last clip holds the last
    ``rf`` frames, a sequence shorter than ``rf`` is padded by repeating its last frame.  With ``kps_left`` /
    ``kps_right`` the flip-TTA twin the callers build first (``main_h3wb.py:268-270``: x negated, left/right
    keypoints swapped) is produced by the same kernel and ``(clips, clips_flip)`` is returned.
    """
    if not inputs_2d.is_cuda:
        raise _native.PafuseError("pafuse_b200.eval_data_prepare needs a CUDA tensor (no CPU fallback)")
    seq = inputs_2d
    if seq.dim() == 4:
        assert seq.shape[0] == 1, "eval_data_prepare takes one sequence"
        seq = seq[0]
    assert seq.dim() == 3 and seq.shape[-1] == 2
    J = seq.shape[1]
    want_flip = kps_left is not None
    perm = None
    if want_flip:
        from .h3wb import flip_permutation
        perm = flip_permutation(kps_left, kps_right, J)
    ctx = _post_context(seq.device, J, int(receptive_field), perm)
    clips, flip = ctx.prepare_clips(seq, want_flip=want_flip)
    return (clips, flip) if want_flip else clips


__all__ = ["wb_pose_from_parts", "project_to_2d", "aggregate_hypotheses", "connection_table", "eval_data_prepare",
           "H3WBSkeleton"]
