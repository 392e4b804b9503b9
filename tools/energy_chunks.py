#!/usr/bin/env python
"""Does running the pass in L2-sized chunks save DRAM energy?  One 640-sequence pass of all three parts with the workspace
chunk (`max_seqs`) set to 640 (default), 64, 32, 16, 10, 8 sequences: with small chunks a kernel's output (tens of MB) is still
in the 126 MB L2 when the next kernel reads it.  Each configuration is replayed from ONE CUDA graph (the chunk loop is inside the
captured pass, so the host launch cost does not matter) for a few seconds while nvidia-smi samples power and clock.

    python tools/energy_chunks.py [seqs] [seconds] [--out profiles/x.json] [--chunks 640,64,32,16,10,8] [--nograph]
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from energy_by_part import Sampler  # noqa: E402
from energy_ablation import measure  # noqa: E402


def main():
    import torch

    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    flags = sys.argv[1:]
    seqs = int(argv[0]) if argv else 640
    seconds = float(argv[1]) if len(argv) > 1 else 4.0
    out_path = flags[flags.index("--out") + 1] if "--out" in flags else ""
    chunks = [int(c) for c in (flags[flags.index("--chunks") + 1] if "--chunks" in flags else "640,64,32,16,10,8").split(",")]
    use_graph = "--nograph" not in flags
    H = 4
    B = seqs // H
    sk = H3WBSkeleton()
    sd = synthetic.synthetic_state_dict(seed=1, depth=8)
    model = pafuse_b200.D3DP(synthetic.default_args(depth=8), sk.joints_left, sk.joints_right, sk, is_train=False,
                             num_proposals=H, sampling_timesteps=1)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    x2d, _ = synthetic.synthetic_inputs(B, seed=1)
    x2d = x2d.cuda()
    x3d = torch.randn(B, H, 27, 134, 3, device="cuda")
    t = torch.full((B,), 999, dtype=torch.long, device="cuda")
    sampler, lines = Sampler(), []
    ref = None
    for chunk in chunks:
        model.max_seqs = chunk
        model._native_dirty = True
        ctx = model.native_context()
        ctx.set_graph_max_seqs(seqs if use_graph else 0)
        out = torch.empty(B, H, 27, 134, 3, device="cuda")
        fn = lambda: model.pred_parts(x2d, x3d, t)
        t0 = time.perf_counter()
        for _ in range(4):                         # the second call with the same pointers captures the graph
            fn()
        r = fn().clone()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        if ref is None:
            ref = r
        same = bool(torch.equal(r, ref))
        ms, w, mhz = measure(sampler, fn, seconds, torch.cuda.synchronize)
        line = {"chunk_seqs": chunk, "sequences": seqs, "graph": use_graph, "graph_replays": ctx.graph_replays(), "ms_per_pass": round(ms, 2),
                "watts_mean": round(w, 1), "sm_mhz_median": mhz, "joules_per_pass": round(w * ms * 1e-3, 1),
                "bit_identical_to_first": same, "setup_s": round(setup, 1)}
        lines.append(line)
        print(json.dumps(line), flush=True)
        time.sleep(1.0)
    if out_path:
        with open(out_path, "w") as f:
            json.dump(lines, f, indent=1)


if __name__ == "__main__":
    main()
