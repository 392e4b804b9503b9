#!/usr/bin/env python
"""Bench lines for the two small / odd workloads of BASELINE.json that bench.py's headline does not cover:

  cfg0  configs[0]: B=2, num_proposals=1, sampling_timesteps=1 (plumbing shape, R=2: launch-bound) -- latency of
        D3DP.forward + wb_pose_from_parts + aggregation with the denoiser pass replayed from a CUDA graph
        (pafuse_set_graph_max_seqs) and without, next to the sum of its kernels' device times;
  cfg5  configs[4]: in_the_wild long-video lifting, T=3000 synthetic OpenPifPaf-style frames -> 112 clips, H=5, K=5,
        through pafuse_b200.in_the_wild.lift_video, writing the reference's .npy output.

    python tools/bench_small.py [--out profiles/r2_small_configs.json]

Prints one JSON object per workload (and writes them to --out).  Needs a B200.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _events(torch, fn, reps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append((a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
    ts.sort()
    return ts[len(ts) // 2]


def cfg0(torch):
    import pafuse_b200
    from pafuse_b200 import synthetic, utils
    from pafuse_b200.h3wb import H3WBSkeleton
    B, H, K = 2, 1, 1
    sk = H3WBSkeleton()
    model = pafuse_b200.D3DP(synthetic.default_args(depth=8), sk.joints_left, sk.joints_right, sk, is_train=False,
                             num_proposals=H, sampling_timesteps=K)
    model.load_state_dict(synthetic.synthetic_state_dict(seed=1, depth=8), strict=False)
    model = model.cuda().eval()
    x2d, x2df = (t.cuda() for t in synthetic.synthetic_inputs(B, seed=1))
    traj, cam = synthetic.synthetic_trajectory(B, seed=1).cuda(), synthetic.h36m_cam0_intrinsics().cuda()
    noises = [n.cuda() for n in synthetic.synthetic_noise(B, H, K, seed=1)]
    model.noise_source = lambda k, shape, device: noises[k]
    ctx = model.native_context()

    def lift():
        out = model(x2d, None, input_2d_flip=x2df)
        wb = pafuse_b200.wb_pose_from_parts(out, sk, mutate_input=False)
        return pafuse_b200.aggregate_hypotheses(wb, traj, cam, x2d)

    res = {}
    for name, max_seqs in (("graph", 96), ("direct", 0)):
        ctx.set_graph_max_seqs(max_seqs)
        r0 = ctx.graph_replays()
        dev_ms, wall_ms = _events(torch, lift, 30, 5)
        res[name] = {"device_ms": dev_ms, "wall_ms": wall_ms, "graph_replays": ctx.graph_replays() - r0}
    a = lift()[0].clone()
    ctx.set_graph_max_seqs(96)
    lift()
    lift()
    same = bool(torch.equal(lift()[0], a))
    # sum of the kernels' own device times (per-launch events; profiling turns the graph replay off)
    post = utils._post_context(torch.device("cuda", torch.cuda.current_device()))
    ctx.profile_enable(True)
    post.profile_enable(True)
    lift()
    prof, prof_post = ctx.profile_read(), post.profile_read()
    ctx.profile_enable(False)
    post.profile_enable(False)
    kernel_ms = sum(v[0] for v in prof.values()) + sum(v[0] for v in prof_post.values())
    launches = sum(v[2] for v in prof.values()) + sum(v[2] for v in prof_post.values())
    return {"workload": "BASELINE configs[0]: B=2 clips, num_proposals=1, sampling_timesteps=1, depth 8, flip-TTA (R=2, 4 sequences per pass)",
            "frames": B * 27, "latency_ms_graph": res["graph"], "latency_ms_direct": res["direct"],
            "sum_of_kernel_times_ms": kernel_ms, "launches_per_lift": launches,
            "latency_over_kernel_sum": res["graph"]["device_ms"] / kernel_ms if kernel_ms else None,
            "frames_per_s_graph": B * 27 / (res["graph"]["wall_ms"] * 1e-3),
            "frames_per_s_direct": B * 27 / (res["direct"]["wall_ms"] * 1e-3),
            "graph_equals_direct_bitwise": same}


def cfg5(torch, T=3000):
    import numpy as np

    import pafuse_b200
    from pafuse_b200 import in_the_wild, synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    H, K = 5, 5
    sk = H3WBSkeleton()
    model = pafuse_b200.D3DP(synthetic.default_args(depth=8), sk.joints_left, sk.joints_right, sk, is_train=False,
                             num_proposals=H, sampling_timesteps=K)
    model.load_state_dict(synthetic.synthetic_state_dict(seed=1, depth=8), strict=False)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(11)
    w, h = 1920, 1080
    det = (torch.rand(T, 133, 3, generator=g) * torch.tensor([w, h, 1.0])).pin_memory()
    out_dir = tempfile.mkdtemp(prefix="pafuse_cfg5_")

    def run(save):
        kp = in_the_wild.keypoints_from_openpifpaf(det, w, h)
        return in_the_wild.lift_video(model, sk, kp, receptive_field=27, bs=1024, video_name="synthetic" if save else None,
                                      out_dir=out_dir)

    dev_ms, wall_ms = _events(torch, lambda: run(False), 3, 1)
    t0 = time.perf_counter()
    res = run(True)
    torch.cuda.synchronize()
    wall_save = (time.perf_counter() - t0) * 1e3
    arr = np.load(res["path"])
    ok = arr.shape == (K, H, T, 134, 3) and arr.dtype == np.float32 and bool(np.isfinite(arr).all())
    size = os.path.getsize(res["path"])
    os.remove(res["path"])
    return {"workload": f"BASELINE configs[4]: in-the-wild long video, T={T} synthetic OpenPifPaf-style frames -> {(T + 26) // 27} clips, "
                        "H=5, K=5, keypoints_from_openpifpaf -> eval_data_prepare (+flip) -> D3DP -> wb_pose_from_parts -> stitch",
            "frames": T, "clips": (T + 26) // 27, "device_ms": dev_ms, "wall_ms": wall_ms, "frames_per_s": T / (wall_ms * 1e-3),
            "with_npy_write": {"wall_ms": wall_save, "frames_per_s": T / (wall_save * 1e-3), "npy_bytes": size,
                               "npy_shape": list(arr.shape), "format_ok": ok,
                               "name": "test_3d_<video>_output.npy (in_the_wild/h3wb_diffusion.py:121,136)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    lines = [cfg0(torch), cfg5(torch)]
    for l in lines:
        print(json.dumps(l))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(lines, f, indent=1)


if __name__ == "__main__":
    main()
