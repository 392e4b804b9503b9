"""GT-dependent multi-hypothesis MPJPE protocols of the reference's ``evaluate()`` on the device.

Same names and argument meaning as ``common/loss.py`` for every call ``main_h3wb.py:344-362`` makes; the whole-body
protocols come out of one kernel pass (``pafuse_mpjpe_metrics``), the part-based ones out of a second
(``pafuse_mpjpe_metrics_parts``):

    mpjpe_diffusion_all_min(pred, target)                           J-Best              loss.py:53-66
    mpjpe_diffusion_all_min(pred, target, mean_pos=True)            P-Agg               loss.py:68-76
    mpjpe_diffusion_reproj(pred, target, reproj, x2d)               J-Agg               loss.py:90-112
    mpjpe_diffusion(pred, target)                                   P-Best              loss.py:114-146
    mpjpe_diffusion(pred, target, part_based=True, dataset=ds)      P-Best part-based   loss.py:119-127,146-154 (+ per part)
    mpjpe_diffusion_all_min(pred, target, mean_pos=True,
                            part_based=True, dataset=ds)            P-Agg part-based    loss.py:41-51,68-86 (+ per part)

``pred`` (B,K,H,F,J,3), ``target`` (B,F,J,3); results are (K,) fp32 tensors like the reference's (the device sums are
fp64).  ``Evaluator`` reproduces the accumulation, weighting and log lines of ``evaluate()`` (main_h3wb.py:364-379,
406-529).  The Procrustes metrics (``args.ft2d.p2``, SVD on the host in the reference) are not implemented.
"""
from __future__ import annotations

import torch

from . import _native
from .utils import _post_context

__all__ = ["evaluate_metrics", "mpjpe_diffusion_all_min", "mpjpe_diffusion_reproj", "mpjpe_diffusion", "Evaluator"]


def _empty(K, device):
    return torch.full((K,), float("nan"), dtype=torch.float32, device=device)


def _means(pred, target, traj=None, cam=None, x2d=None, reproj=None):
    """(K, 3+H) means over (b,f,j): J-Best, P-Agg, J-Agg, per-hypothesis root-centred error."""
    if not pred.is_cuda:
        raise _native.PafuseError("pafuse_b200.loss needs CUDA tensors (no CPU fallback)")
    B, K, H, F, J, _ = pred.shape
    if B == 0:                                                 # torch.mean of an empty tensor: nan, like the reference
        return torch.full((K, 3 + H), float("nan"), dtype=torch.float64, device=pred.device)
    ctx = _post_context(pred.device, J, F)
    return ctx.mpjpe_metrics(pred, target, traj, cam, x2d, reproj) / float(B * F * J)


def _part_tables(dataset, num_kps):
    names = list(dataset.parts_joint_indices.keys())
    part_of, root_of = [-1] * num_kps, [0] * num_kps
    for pi, name in enumerate(names):
        for j in dataset.parts_joint_indices[name]:
            part_of[j] = pi
            root_of[j] = int(dataset.root_indices[name])
    return names, part_of, root_of


def _part_means(pred, target, dataset):
    """names, counts (n_parts,), sums (K, H+1, n_parts) of the part-centred errors (row H: the mean pose)."""
    if not pred.is_cuda:
        raise _native.PafuseError("pafuse_b200.loss needs CUDA tensors (no CPU fallback)")
    assert dataset is not None
    B, K, H, F, J, _ = pred.shape
    names, part_of, root_of = _part_tables(dataset, J)
    counts = torch.tensor([len(dataset.parts_joint_indices[n]) for n in names], dtype=torch.float64, device=pred.device)
    sums = _post_context(pred.device, J, F).mpjpe_metrics_parts(pred, target, part_of, root_of, len(names))
    return names, counts, sums


def evaluate_metrics(pred, target, inputs_traj, cam, inputs_2d):
    """All four whole-body protocols of ``main_h3wb.py:336-349`` (reprojection of ``pred + traj`` included): dict of (K,)."""
    m = _means(pred, target, inputs_traj, cam, inputs_2d)
    return {"J-Best": m[:, 0].float(), "P-Agg": m[:, 1].float(), "J-Agg": m[:, 2].float(),
            "P-Best": m[:, 3:].min(dim=1).values.float()}


def mpjpe_diffusion_all_min(predicted, target, mean_pos=False, part_based=False, dataset=None):
    if not part_based:
        m = _means(predicted, target)
        return m[:, 1].float() if mean_pos else m[:, 0].float()
    if not mean_pos:
        raise NotImplementedError("mpjpe_diffusion_all_min(part_based=True, mean_pos=False) is not called by evaluate()")
    B, K, H, F, J, _ = predicted.shape
    if B == 0:
        return _empty(K, predicted.device), {n: _empty(K, predicted.device) for n in dataset.parts_joint_indices}
    names, counts, sums = _part_means(predicted, target, dataset)
    row = sums[:, H, :]                                        # (K, n_parts) sums of |mean_h pc - gc|
    errors = (row.sum(dim=1) / float(B * F * J)).float()       # joints outside every part contribute 0, like zeros_like
    parts = {n: (row[:, i] / (float(B * F) * counts[i])).float() for i, n in enumerate(names)}
    return errors, parts


def mpjpe_diffusion_reproj(predicted, target, reproj_2d, target_2d):
    return _means(predicted, target, x2d=target_2d, reproj=reproj_2d)[:, 2].float()


def mpjpe_diffusion(predicted, target, mean_pos=False, part_based=False, dataset=None):
    if mean_pos:
        raise NotImplementedError("mpjpe_diffusion(mean_pos=True) is not called by evaluate(); use mpjpe_diffusion_all_min")
    if not part_based:
        m = _means(predicted, target)
        return m[:, 3:].min(dim=1).values.float(), {}
    B, K, H, F, J, _ = predicted.shape
    if B == 0:
        return _empty(K, predicted.device), {n: _empty(K, predicted.device) for n in dataset.parts_joint_indices}
    names, counts, sums = _part_means(predicted, target, dataset)
    per_h = sums[:, :H, :]                                     # (K, H, n_parts)
    total = (per_h.sum(dim=2) / float(B * F * J)).float()      # (K, H) fp32 like the reference's means, so ties break alike
    min_errors, min_inds = total.min(dim=1)
    pick = per_h.gather(1, min_inds.view(K, 1, 1).expand(K, 1, len(names))).squeeze(1)      # (K, n_parts)
    parts = {n: (pick[:, i] / (float(B * F) * counts[i])).float() for i, n in enumerate(names)}
    return min_errors, parts


class Evaluator:
    """Accumulation and report of ``evaluate()`` (main_h3wb.py:209-224, 364-379, 406-529) for one action / one run.

    ``update`` takes what the reference has at hand after one sub-batch (main_h3wb.py:322-342): whole-body predictions
    ``(b,K,H,F,J,3)``, the whole-body target ``(b,F,J,3)``, the root trajectory ``(b,F,1,3)``, the camera intrinsics and
    the 2D input.  Every protocol is weighted by ``b*F`` (``batch_multiplier``) like there; P-Best picks its hypothesis
    per sub-batch, also like there.  ``results()`` returns mm values, ``log_lines()`` the text of the log file.
    """

    PROTOCOLS = ("J_Best", "P_Best", "P_Agg", "J_Agg", "P_Best_pb", "P_Agg_pb")

    def __init__(self, dataset, sampling_timesteps, device="cuda"):
        self.dataset, self.K, self.N = dataset, int(sampling_timesteps), 0
        self.parts = list(dataset.parts_joint_indices.keys())
        z = lambda: torch.zeros(self.K, dtype=torch.float32, device=device)
        self.sums = {p: z() for p in self.PROTOCOLS}
        self.sums.update({f"P_Best_pb_{n}": z() for n in self.parts})
        self.sums.update({f"P_Agg_pb_{n}": z() for n in self.parts})

    def update(self, predicted_wb, target_wb, inputs_traj, cam, inputs_2d):
        b, K, H, F, J, _ = predicted_wb.shape
        assert K == self.K
        if b == 0:
            return
        mult = float(b * F)
        m = evaluate_metrics(predicted_wb, target_wb, inputs_traj, cam, inputs_2d)
        e_h_pb, e_parts = mpjpe_diffusion(predicted_wb, target_wb, part_based=True, dataset=self.dataset)
        e_agg_pb, e_agg_parts = mpjpe_diffusion_all_min(predicted_wb, target_wb, mean_pos=True, part_based=True,
                                                        dataset=self.dataset)
        for key, v in (("J_Best", m["J-Best"]), ("P_Best", m["P-Best"]), ("P_Agg", m["P-Agg"]), ("J_Agg", m["J-Agg"]),
                       ("P_Best_pb", e_h_pb), ("P_Agg_pb", e_agg_pb)):
            self.sums[key] += mult * v
        for n in self.parts:
            self.sums[f"P_Best_pb_{n}"] += mult * e_parts[n]
            self.sums[f"P_Agg_pb_{n}"] += mult * e_agg_parts[n]
        self.N += b * F

    def results(self):
        """{name: (K,) tensor in mm} = ``(epoch_sum / N) * 1000`` (main_h3wb.py:417-432)."""
        n = max(self.N, 1)
        return {k: (v / n) * 1000 for k, v in self.sums.items()}

    def log_lines(self, action=None):
        """The lines ``evaluate()`` writes to ``h36m_test_log_H*_K*.txt`` (main_h3wb.py:409-415, 440-515)."""
        r = {k: v.cpu() for k, v in self.results().items()}
        out = [] if action is None else ["----" + action + "----"]
        hands = lambda pre, ii: (r[pre + "_right_hand"][ii].item() + r[pre + "_left_hand"][ii].item()) / 2.
        for ii in range(self.K):
            p1 = "step %d : Protocol #1 Error (MPJPE) " % ii
            out.append(p1 + "J_Best: %f mm" % r["J_Best"][ii].item())
            out.append(p1 + "P_Agg: %f mm" % r["P_Agg"][ii].item())
            out.append(p1 + "J_Agg: %f mm" % r["J_Agg"][ii].item())
            for title, pre in (("-----------------> Part-Based Evaluation <-----------------", "P_Best"),
                               ("-----------------> Part-Based Evaluation Aggregation <-----------------", "P_Agg")):
                out += [title, title]                          # the reference writes the banner twice (main_h3wb.py:457-461)
                out.append(p1 + "%s Part-Based: %f mm" % (pre, r[pre + "_pb"][ii].item()))
                out.append(p1 + "%s Part-Based BODY: %f mm" % (pre, r[pre + "_pb_body"][ii].item()))
                out.append(p1 + "%s Part-Based FACE: %f mm" % (pre, r[pre + "_pb_face"][ii].item()))
                if "left_hand" in self.parts and "right_hand" in self.parts:
                    out.append(p1 + "%s Part-Based HANDS: %f mm" % (pre, hands(pre + "_pb", ii)))
                    out.append(p1 + "%s Part-Based LEFT HAND: %f mm" % (pre, r[pre + "_pb_left_hand"][ii].item()))
                    out.append(p1 + "%s Part-Based RIGHT HAND: %f mm" % (pre, r[pre + "_pb_right_hand"][ii].item()))
        out.append("----------")
        return out
