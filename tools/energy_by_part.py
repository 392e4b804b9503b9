#!/usr/bin/env python
"""Power, clock and energy of the denoiser passes, per part: the step is bound by the 1000 W cap,
so what a pass COSTS is joules, not only microseconds.

Loops one denoiser pass (MixSTE2.forward of one part, or D3DP.pred_parts of all three) for a few seconds each while
nvidia-smi samples power and SM clock every 100 ms, and prints per configuration: ms per pass, mean W, median SM MHz,
J per pass, and J per algorithmic TFLOP (matmul FLOPs of the fp32 layers).

    python tools/energy_by_part.py [seqs] [seconds] [--out profiles/x.json]
"""
from __future__ import annotations

import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLOP_R1 = {"body": 24_871_514_112, "face": 24_842_567_680, "hands": 19_670_624_256}   # per (clip, hypothesis) forward


class Sampler:
    def __init__(self):
        self.rows, self.proc = [], None

    def start(self):
        self.rows = []
        self.proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits",
                                      "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            try:
                p, c = (float(v) for v in line.split(","))
                self.rows.append((p, c))
            except ValueError:
                pass

    def stop(self):
        self.proc.terminate()
        rows = self.rows[2:] if len(self.rows) > 4 else self.rows      # drop the ramp
        if not rows:
            return None, None
        return statistics.mean(r[0] for r in rows), statistics.median(r[1] for r in rows)


def main():
    import torch

    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton, merged_part_indices
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    seqs = int(argv[0]) if argv else 640
    seconds = float(argv[1]) if len(argv) > 1 else 4.0
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else ""
    H = 4
    B = seqs // H
    sk = H3WBSkeleton()
    parts = merged_part_indices(sk.parts_joint_indices)
    sd = synthetic.synthetic_state_dict(seed=1, depth=8)
    model = pafuse_b200.D3DP(synthetic.default_args(depth=8), sk.joints_left, sk.joints_right, sk, is_train=False,
                             num_proposals=H, sampling_timesteps=1)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    x2d, _ = synthetic.synthetic_inputs(B, seed=1)
    x2d = x2d.cuda()
    x3d = torch.randn(B, H, 27, 134, 3, device="cuda")
    t = torch.full((B,), 999, dtype=torch.long, device="cuda")
    runs = {"all parts": (lambda: model.pred_parts(x2d, x3d, t), sum(FLOP_R1.values()))}
    for name, idx in parts.items():
        C = {"body": 384, "face": 224, "hands": 256}[name]
        net = pafuse_b200.MixSTE2(num_frame=27, num_joints=len(idx), in_chans=5, embed_dim_ratio=C, depth=8, num_heads=8,
                                  mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_path_rate=0, is_train=False)
        net.load_state_dict({k[len(f"pose_estimator.{name}."):]: v for k, v in sd.items() if k.startswith(f"pose_estimator.{name}.")})
        net = net.cuda().eval()
        a, b = x2d[:, :, idx].contiguous(), x3d[:, :, :, idx].contiguous()
        runs[name] = ((lambda net=net, a=a, b=b: net(a, b, t)), FLOP_R1[name])
    sampler = Sampler()
    lines = []
    for name, (fn, flop_r1) in runs.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        sampler.start()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            n += 2
        dt = time.perf_counter() - t0
        watts, mhz = sampler.stop()
        ms = dt / n * 1e3
        tflop = flop_r1 * seqs / 1e12
        line = {"pass": name, "sequences": seqs, "ms_per_pass": round(ms, 2), "watts_mean": None if watts is None else round(watts, 1),
                "sm_mhz_median": mhz, "joules_per_pass": None if watts is None else round(watts * ms * 1e-3, 1),
                "algorithmic_tflop_per_pass": round(tflop, 2),
                "joules_per_algorithmic_tflop": None if watts is None else round(watts * ms * 1e-3 / tflop, 2),
                "algorithmic_tflops": round(tflop / (ms * 1e-3), 1)}
        lines.append(line)
        print(json.dumps(line))
        time.sleep(1.0)
    if out_path:
        with open(out_path, "w") as f:
            json.dump(lines, f, indent=1)


if __name__ == "__main__":
    main()
