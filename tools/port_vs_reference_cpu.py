#!/usr/bin/env python
"""How representative is the CPU arm?  bench.py's `cpu_baseline` / `--impl reference` time the oracle PORT (kind "port":
the Python reference does not exist on the GPU box).  This tool runs, in the build container where /root/reference is
present, the UNMODIFIED reference (`common/diffusionpose.py:D3DP.forward` through oracle/ref_harness.py) and the port on
the SAME sample (clips, H, K, flip-TTA, depth 8, same injected noise, same thread count), alternately, and prints both
times and whether the two results are identical -- so the port's frames/s on the GPU box can be read as the reference's.

    python tools/port_vs_reference_cpu.py [clips] [H] [K] [repeats]      # default 2 5 5 3

Test infrastructure only (imports oracle/); never on the product path.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from oracle import pafuse_oracle as orc
    from oracle import ref_harness
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton, merged_part_indices
    argv = sys.argv[1:]
    clips = int(argv[0]) if argv else 2
    H = int(argv[1]) if len(argv) > 1 else 5
    K = int(argv[2]) if len(argv) > 2 else 5
    reps = int(argv[3]) if len(argv) > 3 else 3
    if not ref_harness.reference_available():
        print(json.dumps({"unavailable": "no /root/reference here"}))
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sk = H3WBSkeleton()
    sd = synthetic.synthetic_state_dict(seed=1, depth=8)
    x2d, x2df = synthetic.synthetic_inputs(clips, seed=1)
    noises = synthetic.synthetic_noise(clips, H, K, seed=1)
    parts = merged_part_indices(sk.parts_joint_indices)
    model, _ = ref_harness.build_reference_model(synthetic.default_args(depth=8), sk, sd, H, K)

    def run_ref():
        t0 = time.perf_counter()
        out = ref_harness.reference_forward(model, x2d, x2df, noises)
        return time.perf_counter() - t0, out

    def run_port():
        t0 = time.perf_counter()
        with torch.no_grad():
            out = orc.ddim_sample_flip(sd, parts, x2d, x2df, noises, sk.joints_left, sk.joints_right, H, K, depth=8)
        return time.perf_counter() - t0, out

    run_ref(), run_port()                                                       # warm-up
    t_ref, t_port, same = [], [], True
    for _ in range(reps):
        a, ref_out = run_ref()
        b, port_out = run_port()
        t_ref.append(a)
        t_port.append(b)
        same = same and torch.equal(ref_out, port_out)
    frames = clips * 27
    line = {"sample": f"{clips} clips x 27 frames, H={H} K={K} flip-TTA depth 8", "threads": threads,
            "reference_s": [round(t, 3) for t in t_ref], "port_s": [round(t, 3) for t in t_port],
            "reference_frames_per_s": round(frames / min(t_ref), 2), "port_frames_per_s": round(frames / min(t_port), 2),
            "port_over_reference_time": round(min(t_port) / min(t_ref), 3), "outputs_identical": bool(same)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
