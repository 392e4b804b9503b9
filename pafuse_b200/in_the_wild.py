"""Long-video lifting, the B200 side of the reference's in-the-wild driver (BASELINE config 5).

Mirrors ``in_the_wild/h3wb_diffusion.py:57-77,105-133`` and ``in_the_wild/utils.py:322-376``: OpenPifPaf
detections -> normalised 134-keypoint 2D input, non-overlapping 27-frame clips with a right-aligned last clip,
sub-batches of ``bs`` clips through ``D3DP`` + ``wb_pose_from_parts``, clips stitched back to the video length, mean
pose over the hypotheses of the last step (``in_the_wild/visualization.py:253-255``).  JSON parsing stays on the
host; every tensor operation is a kernel of the C-ABI library (no CPU fallback).
"""
from __future__ import annotations

import json

import torch

from . import _native
from .h3wb import flip_permutation
from .utils import _post_context, eval_data_prepare, wb_pose_from_parts

__all__ = ["read_openpifpaf_json", "keypoints_from_openpifpaf", "evaluate_diffusion", "stitch_predictions", "lift_video",
           "save_prediction"]


def read_openpifpaf_json(path_or_lines):
    """JSON-lines file written by openpifpaf (one frame per line) -> fp32 tensor ``(T, 133, 3)`` of the first
    prediction's ``(x, y, confidence)`` triples (``h3wb_diffusion.py:57-66``)."""
    if isinstance(path_or_lines, str):
        with open(path_or_lines, "r") as f:
            lines = [l for l in f if l.strip()]
    else:
        lines = list(path_or_lines)
    frames = []
    for line in lines:
        kp = json.loads(line) if isinstance(line, str) else line
        frames.append(kp["predictions"][0]["keypoints"])
    return torch.tensor(frames, dtype=torch.float32).reshape(len(frames), -1, 3)


def keypoints_from_openpifpaf(detections, width, height, device="cuda"):
    """``(T,133,3)`` pixel detections -> ``(T,134,2)`` model input: joint 0 = mean of joints 12 and 13
    (``h3wb_diffusion.py:64-69``), then ``normalize_screen_coordinates`` (``common/camera.py:7-11``)."""
    det = detections.to(device=device, dtype=torch.float32)
    if not det.is_cuda:
        raise _native.PafuseError("pafuse_b200.in_the_wild needs a CUDA device (no CPU fallback)")
    J = det.shape[1] + 1
    return _post_context(det.device, J).keypoints_from_detections(det, width, height)


def evaluate_diffusion(model_pos, dataset, keypoints, receptive_field=27, bs=1024):
    """``in_the_wild/utils.py:322-376``: keypoints ``(T,134,2)`` (CUDA) -> whole-body predictions
    ``(N,K,H,rf,134,3)`` for the ``N = ceil(T/rf)`` clips, on the device."""
    sym = dataset.keypoints_metadata["keypoints_symmetry"]
    kps_left, kps_right = list(sym[0]), list(sym[1])
    inputs_2d, inputs_2d_flip = eval_data_prepare(receptive_field, keypoints, kps_left, kps_right)
    outs = []
    with torch.no_grad():
        for b0 in range(0, inputs_2d.shape[0], bs):
            pred = model_pos(inputs_2d[b0:b0 + bs], None, input_2d_flip=inputs_2d_flip[b0:b0 + bs])
            outs.append(wb_pose_from_parts(pred, dataset=dataset))
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


def stitch_predictions(prediction, total_frame):
    """``h3wb_diffusion.py:119-133``: ``(N,K,H,rf,J,3)`` -> ``(K,H,T,J,3)``."""
    J, rf = prediction.shape[-2], prediction.shape[-3]
    return _post_context(prediction.device, J, rf).stitch_clips(prediction, int(total_frame))


def save_prediction(prediction, video_name, out_dir="outputs"):
    """The reference's on-disk result (``h3wb_diffusion.py:121,136``): fp32 array ``(K,H,T,134,3)`` (camera frame,
    before the world-frame post-processing) saved as ``<out_dir>/<video_name>/test_3d_<video_name>_output.npy``.
    Returns the path."""
    import os

    import numpy as np
    d = os.path.join(out_dir, video_name)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"test_3d_{video_name}_output.npy")
    np.save(path, prediction.detach().to("cpu", torch.float32).numpy(), allow_pickle=True)
    return path


def lift_video(model_pos, dataset, keypoints, receptive_field=27, bs=1024, video_name=None, out_dir="outputs"):
    """Whole driver: returns ``{"prediction": (K,H,T,134,3), "mean_pose": (T,134,3)}`` where ``mean_pose`` is the mean
    over the hypotheses of the last sampling step; with ``video_name`` the prediction is also written in the
    reference's ``.npy`` format (``"path"`` in the result)."""
    T = keypoints.shape[0]
    pred = evaluate_diffusion(model_pos, dataset, keypoints, receptive_field, bs)
    out = stitch_predictions(pred, T)
    res = {"prediction": out, "mean_pose": out[-1].mean(dim=0)}
    if video_name is not None:
        res["path"] = save_prediction(out, video_name, out_dir)
    return res
