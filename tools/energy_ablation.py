#!/usr/bin/env python
"""Where do the joules of a denoiser pass go?  Needs the ablation build (tools/build_variant.sh ablate -DPAFUSE_ABLATE) loaded
through PAFUSE_LIB: its GEMM and attention kernels read PAFUSE_ABLATE at every launch and
    1  GEMM epilogue warps hand the accumulator straight back (no tensor-memory loads, math, staging, output stores)
    2  GEMM issuer skips its MMAs (operands still travel DRAM -> shared memory, barriers still cycle)
    4  attention softmax / output warps skip their work (same barrier protocol)
    8  attention issuer skips its MMAs
Skipped work leaves the previous contents of the buffers in place (real data from the full passes run first), so operand
toggling stays representative.  Every configuration loops one 640-sequence pass of all three parts for a few seconds while
nvidia-smi samples power and SM clock; printed: ms per pass, mean W, J per pass.  Calibration lines: idle (context alive, no work),
a DRAM copy (torch `copy_`, bytes/J) and a cuBLAS bf16 GEMM (J per issued TFLOP) under the same cap.

    PAFUSE_LIB=_ab_ablate/libpafuse_b200.so python tools/energy_ablation.py [seqs] [seconds] [--out profiles/x.json]

The numbers are never bench values (work is skipped); the default build has none of this code.
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


def measure(sampler, fn, seconds, sync):
    for _ in range(2):
        fn()
    sync()
    sampler.start()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(2):
            fn()
        sync()
        n += 2
    dt = time.perf_counter() - t0
    watts, mhz = sampler.stop()
    return dt / n * 1e3, watts, mhz


def main():
    import torch

    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    seqs = int(argv[0]) if argv else 640
    seconds = float(argv[1]) if len(argv) > 1 else 4.0
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else ""
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
    sync = torch.cuda.synchronize

    def emit(line):
        lines.append(line)
        print(json.dumps(line), flush=True)

    # calibration
    sampler.start()
    time.sleep(3.0)
    w, mhz = sampler.stop()
    emit({"run": "idle (context alive)", "watts_mean": w and round(w, 1), "sm_mhz_median": mhz})
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    ms, w, mhz = measure(sampler, lambda: b.copy_(a), seconds, sync)
    gb = 2 * a.numel() * 2 / 1e9
    emit({"run": "DRAM copy 2 GiB -> 2 GiB", "ms": round(ms, 3), "gbs": round(gb / ms * 1e3, 1), "watts_mean": round(w, 1), "sm_mhz_median": mhz,
          "joules_per_tb": round(w * ms * 1e-3 / (gb / 1e3), 1)})
    del a, b
    m1 = torch.randn(8192, 8192, device="cuda").bfloat16()
    m2 = torch.randn(8192, 8192, device="cuda").bfloat16()
    ms, w, mhz = measure(sampler, lambda: torch.matmul(m1, m2), seconds, sync)
    tf = 2 * 8192 ** 3 / 1e12
    emit({"run": "cuBLAS bf16 8192^3", "ms": round(ms, 3), "tflops": round(tf / ms * 1e3, 1), "watts_mean": round(w, 1), "sm_mhz_median": mhz,
          "joules_per_issued_tflop": round(w * ms * 1e-3 / tf, 3)})
    del m1, m2
    time.sleep(1.0)

    fn = lambda: model.pred_parts(x2d, x3d, t)
    names = {0: "full pass", 1: "GEMM epilogues off", 2: "GEMM MMAs off", 3: "GEMM epilogues + MMAs off (operand loads only)",
             4: "attention softmax / output off", 8: "attention MMAs off", 12: "attention softmax + MMAs off (operand loads only)",
             5: "all epilogue / softmax work off", 10: "all MMAs off", 15: "everything off but the operand loads (+ LayerNorm, embed, head kernels)"}
    for mask in (0, 1, 2, 3, 4, 8, 12, 5, 10, 15, 0):
        os.environ["PAFUSE_ABLATE"] = str(mask)
        ms, w, mhz = measure(sampler, fn, seconds, sync)
        emit({"run": names[mask], "mask": mask, "sequences": seqs, "ms_per_pass": round(ms, 2), "watts_mean": round(w, 1), "sm_mhz_median": mhz,
              "joules_per_pass": round(w * ms * 1e-3, 1)})
        time.sleep(1.0)
    os.environ["PAFUSE_ABLATE"] = "0"
    if out_path:
        with open(out_path, "w") as f:
            json.dump(lines, f, indent=1)


if __name__ == "__main__":
    main()
