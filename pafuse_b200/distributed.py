"""Multi-GPU lifting: one process per GPU, clips or hypotheses sharded, one collective at the end.

The reference's only multi-GPU mechanism is ``nn.DataParallel`` over clips
(``main_h3wb.py:699-705``), which re-broadcasts the weights on every forward and
gathers on GPU 0.  Here every rank holds the weights once and owns a contiguous
range of clips (``mode='clips'``, BASELINE config 4) or of hypotheses
(``mode='hypotheses'``, config 3).  Nothing is exchanged inside the K-step DDIM loop
(hypotheses never interact inside the denoiser, ``mixste.py:227-230``); the only
data-path collectives are

  * clips:       one all-gather of the aggregated poses ``(B/G,K,F,J,3)`` x2,
  * hypotheses:  one all-to-all that turns the hypothesis shards ``(B,K,H/G,F,J,3)`` into
                 clip shards ``(B/G,K,H,F,J,3)`` (rank order == hypothesis order, so the
                 aggregation kernel sees exactly the single-GPU tensor), the local
                 aggregation, and the same all-gather.

Noise parity: the sampler's draws are a pure function of (seed, draw number, global
element index) -- a counter-based Philox generator inside the library (``PhiloxNoise`` ->
``pafuse_randn``) -- so every rank produces exactly ITS slice of the global ``(B,H,F,J,3)``
tensors and a G-GPU run is bit-identical per element to the 1-GPU run
without any rank drawing more than its share.  ``ShardedNoise`` (every rank draws the global
tensor from an identically seeded torch generator and keeps its slice) remains for the
CPU/gloo tests of the sharding logic and as ``noise="torch"``.

The compute is behind a small ``engine`` interface so the sharding logic is testable
on CPU with the ``gloo`` backend (tests plug the oracle in; the product engine is
``CudaEngine`` and has no CPU path).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
import torch.distributed as dist

__all__ = ["shard_range", "ShardedNoise", "PhiloxNoise", "CudaEngine", "LiftResult", "lift", "lift_sharded"]


def shard_range(n: int, world: int, rank: int):
    """Contiguous balanced partition of range(n): the first ``n % world`` ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class ShardedNoise:
    """``noise_source`` for ``D3DP``: draws the global tensor, returns the local slice.

    Draw k of the sampler has the global shape ``(B,H,F,J,3)`` (``diffusionpose.py:283,308``).
    """

    def __init__(self, seed: int, global_B: int, global_H: int, b_range, h_range, device):
        self.global_B, self.global_H = global_B, global_H
        self.b_range, self.h_range = b_range, h_range
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)

    def __call__(self, k, shape, device):
        full = torch.randn((self.global_B, self.global_H) + tuple(shape[2:]), generator=self.gen,
                           device=self.gen.device, dtype=torch.float32)
        (b0, b1), (h0, h1) = self.b_range, self.h_range
        assert (b1 - b0, h1 - h0) == tuple(shape[:2]), "local shape does not match the shard"
        return full[b0:b1, h0:h1].contiguous().to(device)


class PhiloxNoise:
    """``noise_source`` for ``D3DP`` on CUDA: the local slice of the global draw, generated in place.

    Element ``(b,h,f,j,c)`` of draw ``k`` is the value of the library's counter-based generator at index
    ``((b*H + h)*F*J*3 + ...)`` of stream ``k`` under ``seed`` (``pafuse_randn``): a clip shard is one contiguous
    run, a hypothesis shard is ``B`` runs of ``(h1-h0)*F*J*3`` elements with stride ``H*F*J*3``.
    """

    def __init__(self, seed: int, global_B: int, global_H: int, b_range, h_range, ctx):
        self.seed, self.global_B, self.global_H = int(seed), global_B, global_H
        self.b_range, self.h_range, self.ctx = b_range, h_range, ctx

    def __call__(self, k, shape, device):
        (b0, b1), (h0, h1) = self.b_range, self.h_range
        assert (b1 - b0, h1 - h0) == tuple(shape[:2]), "local shape does not match the shard"
        per = 1
        for d in shape[2:]:
            per *= int(d)
        H = self.global_H
        out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
        if (h0, h1) == (0, H):                                 # clips: one contiguous run
            self.ctx.randn(self.seed, k, b0 * H * per, 1, (b1 - b0) * H * per, (b1 - b0) * H * per, out=out)
        else:
            self.ctx.randn(self.seed, k, (b0 * H + h0) * per, b1 - b0, (h1 - h0) * per, H * per, out=out)
        return out


class LiftResult(NamedTuple):
    jagg: torch.Tensor                 # (B,K,F,J,3)  J-Agg pose
    pagg: torch.Tensor                 # (B,K,F,J,3)  P-Agg pose
    select: Optional[torch.Tensor]     # (B,K,F,J) int32 hypothesis chosen by J-Agg
    pred: Optional[torch.Tensor]       # (B,K,H,F,J,3) whole-body hypotheses (when kept)


class CudaEngine:
    """The product engine: ``D3DP`` + the post-processing kernels, all through the C ABI."""

    def __init__(self, model, dataset):
        self.model, self.dataset = model, dataset

    def sample(self, x2d, x2d_flip, num_proposals, noise_source):
        m = self.model
        prev = (m.num_proposals, m.noise_source)
        m.num_proposals, m.noise_source = num_proposals, noise_source
        try:
            return m(x2d, None, input_2d_flip=x2d_flip)
        finally:
            m.num_proposals, m.noise_source = prev

    def reassemble(self, pred):
        from .utils import wb_pose_from_parts
        return wb_pose_from_parts(pred, self.dataset, mutate_input=False)

    def aggregate(self, wb, traj, cam, x2d):
        from .utils import aggregate_hypotheses
        return aggregate_hypotheses(wb, traj, cam, x2d, return_select=True)

    def noise(self, seed, B, H, b_range, h_range, device):
        """The sampler's draws for this shard (counter-based, generated on the device by the library)."""
        return PhiloxNoise(seed, B, H, b_range, h_range, self.model.native_context(device))


def lift(engine, x2d, x2d_flip, traj, cam, num_proposals, noise_source=None, keep_hypotheses=False) -> LiftResult:
    """Single-device lift: sampler -> part re-assembly -> J-Agg / P-Agg (main_h3wb.py:322-362)."""
    pred = engine.sample(x2d, x2d_flip, num_proposals, noise_source)
    wb = engine.reassemble(pred)
    jagg, pagg, sel = engine.aggregate(wb, traj, cam, x2d)
    return LiftResult(jagg, pagg, sel, wb if keep_hypotheses else None)


def _all_gather_rows(t: torch.Tensor, counts, group=None) -> torch.Tensor:
    """Concatenate dim-0 shards of unequal length (padded to the longest for one fixed-size all-gather)."""
    world = len(counts)
    if world == 1:
        return t
    longest = max(counts)
    buf = t
    if t.shape[0] < longest:
        buf = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
    out = torch.empty((world * longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if all(c == longest for c in counts):
        return out
    return torch.cat([out[r * longest: r * longest + c] for r, c in enumerate(counts)], dim=0)


def _all_gather_packed(jagg, pagg, sel, counts, group=None):
    """ONE all-gather for the three results of a step: [jagg | pagg | select(int32 bits)] of every rank, each padded
    to the longest shard.  (Three blocking collectives per step were three rendez-vous with the slowest rank.)"""
    world, longest = len(counts), max(counts)
    n, per = jagg.shape[0], jagg[0].numel() if jagg.shape[0] else int(torch.tensor(jagg.shape[1:]).prod())
    per_sel = per // 3
    width = longest * (2 * per + per_sel)
    buf = torch.zeros(width, dtype=torch.float32, device=jagg.device)
    buf[: n * per] = jagg.reshape(-1)
    buf[longest * per: longest * per + n * per] = pagg.reshape(-1)
    buf[2 * longest * per: 2 * longest * per + n * per_sel] = sel.reshape(-1).to(torch.int32).view(torch.float32)
    out = torch.empty(world * width, dtype=torch.float32, device=jagg.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, width)
    tail = tuple(jagg.shape[1:])
    js = [out[r, : c * per].reshape((c,) + tail) for r, c in enumerate(counts)]
    ps = [out[r, longest * per: longest * per + c * per].reshape((c,) + tail) for r, c in enumerate(counts)]
    ss = [out[r, 2 * longest * per: 2 * longest * per + c * per_sel].view(torch.int32).reshape((c,) + tail[:-1])
          for r, c in enumerate(counts)]
    return torch.cat(js), torch.cat(ps), torch.cat(ss)


def lift_sharded(engine, x2d, x2d_flip, traj, cam, num_proposals, mode="clips", seed=0, rank=None, world=None,
                 group=None, gather=True, noise_device=None, noise=None) -> LiftResult:
    """Lift the GLOBAL batch ``x2d (B,F,J,2)`` with this rank's shard of the work.

    Every rank passes the same global inputs (they are tiny: 7 KB per clip) and gets, when
    ``gather`` is true, the full ``(B,K,F,J,3)`` aggregated poses; with ``gather=False`` only
    its clip shard (rows ``shard_range(B, world, rank)``).
    """
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    B, H = x2d.shape[0], num_proposals
    device = x2d.device
    noise_device = noise_device or device
    if noise is None:
        noise = "philox" if hasattr(engine, "noise") and device.type == "cuda" else "torch"

    def make_noise(b_range, h_range):
        if noise == "philox":
            return engine.noise(seed, B, H, b_range, h_range, device)
        return ShardedNoise(seed, B, H, b_range, h_range, noise_device)
    cam_of = (lambda b0, b1: cam[b0:b1]) if (cam.dim() == 2 and cam.shape[0] == B and B > 1) else (lambda b0, b1: cam)
    b0, b1 = shard_range(B, world, rank)
    clip_counts = [shard_range(B, world, r)[1] - shard_range(B, world, r)[0] for r in range(world)]

    if mode == "clips":
        res = lift(engine, x2d[b0:b1].contiguous(), None if x2d_flip is None else x2d_flip[b0:b1].contiguous(),
                   None if traj is None else traj[b0:b1].contiguous(), cam_of(b0, b1), H, make_noise((b0, b1), (0, H)))
        jagg, pagg, sel = res.jagg, res.pagg, res.select
    elif mode == "hypotheses":
        h0, h1 = shard_range(H, world, rank)
        if h1 == h0:
            raise ValueError(f"num_proposals={H} < world size {world}: nothing to do on rank {rank}")
        wb = engine.reassemble(engine.sample(x2d, x2d_flip, h1 - h0, make_noise((0, B), (h0, h1))))   # (B,K,h,F,J,3)
        if world > 1:
            # all-to-all: send clip block q of my hypotheses to rank q, receive my clip block of everyone's
            K = wb.shape[1]
            tail = tuple(wb.shape[3:])
            h_counts = [shard_range(H, world, r)[1] - shard_range(H, world, r)[0] for r in range(world)]
            send = torch.cat([wb[shard_range(B, world, q)[0]: shard_range(B, world, q)[1]].reshape(-1)
                              for q in range(world)])
            per = K * int(torch.tensor(tail).prod())
            in_splits = [clip_counts[q] * (h1 - h0) * per for q in range(world)]
            out_splits = [(b1 - b0) * h_counts[r] * per for r in range(world)]
            recv = torch.empty(sum(out_splits), dtype=wb.dtype, device=device)
            dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
            blocks, off = [], 0
            for r in range(world):
                blocks.append(recv[off: off + out_splits[r]].reshape((b1 - b0, K, h_counts[r]) + tail))
                off += out_splits[r]
            wb = torch.cat(blocks, dim=2).contiguous()                               # (B/G,K,H,F,J,3), rank order = h order
        else:
            wb = wb[b0:b1]
        jagg, pagg, sel = engine.aggregate(wb, None if traj is None else traj[b0:b1].contiguous(), cam_of(b0, b1),
                                           x2d[b0:b1].contiguous())
    else:
        raise ValueError(f"unknown sharding mode {mode!r}")

    if gather and world > 1:
        if sel is not None:
            jagg, pagg, sel = _all_gather_packed(jagg, pagg, sel, clip_counts, group)
        else:
            jagg = _all_gather_rows(jagg, clip_counts, group)
            pagg = _all_gather_rows(pagg, clip_counts, group)
    return LiftResult(jagg, pagg, sel, None)
