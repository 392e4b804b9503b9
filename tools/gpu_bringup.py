"""GPU bring-up ladder: each rung runs in its own subprocess with a timeout so a
trap or hang in one kernel cannot take the whole gpurun call down.

    python tools/gpu_bringup.py [rung ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RUNGS = {}


def rung(f):
    RUNGS[f.__name__] = f
    return f


def _ctx(num_kps=134):
    import torch
    from pafuse_b200 import _native
    return _native.NativeContext(27, num_kps, 1, 8, [32], [[0]], list(range(num_kps)), 1.0, 1, torch.device("cuda", 0))


@rung
def create():
    import torch
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    c = _ctx()
    print("version", c.lib.pafuse_version().decode())


def _linear_case(M, N, K, epi, simt):
    import torch
    torch.manual_seed(0)
    c = _ctx()
    x = torch.randn(M, K, device="cuda")
    w = (torch.rand(N, K, device="cuda") * 2 - 1) / K ** 0.5
    b = torch.randn(N, device="cuda") * 0.1
    ref = (x.double() @ w.double().t() + b.double())
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    y = None
    if epi == 2:
        y0 = torch.randn(M, N, device="cuda")
        ref = ref + y0.double()
        y = y0.clone()
    out = c.linear(x, w, b, epilogue=epi, use_simt=simt, y=y)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs()
    scale = ref.abs().mean().item()
    print(f"linear M={M} N={N} K={K} epi={epi} simt={simt}: max abs err {err.max().item():.3e} "
          f"mean {err.mean().item():.3e} (ref mean abs {scale:.3f})")
    return err.max().item()


@rung
def linear_simt():
    assert _linear_case(300, 224, 224, 0, True) < 1e-5


def _cg(n):
    _ctx().set_gemm_cta_group(n)


@rung
def linear_cg1_small():
    _cg(1)
    assert _linear_case(128, 256, 64, 0, False) < 1e-5
    assert _linear_case(300, 224, 224, 0, False) < 1e-5


@rung
def linear_tc_small():
    _cg(2)
    assert _linear_case(128, 256, 64, 0, False) < 1e-5


@rung
def linear_tc_k():
    _cg(2)
    assert _linear_case(256, 256, 256, 0, False) < 1e-5
    assert _linear_case(300, 224, 224, 0, False) < 1e-5
    assert _linear_case(1000, 1152, 384, 0, False) < 1e-5


@rung
def linear_tc_shapes():
    for (N, K) in [(1152, 384), (384, 384), (768, 384), (384, 768), (672, 224), (224, 224), (448, 224), (224, 448),
                   (768, 256), (256, 256), (512, 256), (256, 512)]:
        for epi in (0, 1, 2):
            assert _linear_case(1000, N, K, epi, False) < 1e-5


@rung
def linear_tc_big():
    import torch
    assert _linear_case(148 * 128 * 3 + 77, 1152, 384, 0, False) < 1e-5
    # throughput probe
    c = _ctx()
    from pafuse_b200 import _native
    M, N, K = 414720, 1152, 384
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.zeros(N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    for _ in range(2):
        c.linear(x, w, b, y=y)
    torch.cuda.synchronize()
    t0 = time.time()
    c.linear(x, w, b, y=y)
    torch.cuda.synchronize()
    print(f"pafuse_linear wall (incl. split + alloc) {1e3 * (time.time() - t0):.2f} ms; "
          f"algorithmic {2 * M * N * K / 1e12:.2f} TFLOP")


def _attention(simt, cases, S=3):
    import torch
    torch.manual_seed(0)
    c = _ctx()
    c.set_debug_simt_attention(simt)
    for (J, C) in cases:
        for temporal in (False, True):
            F = 27
            qkv = torch.randn(S * F * J, 3 * C, device="cuda")
            out = c.attention(qkv, S, J, C, temporal)
            torch.cuda.synchronize()
            hd = C // 8
            t = qkv.double().reshape(S, F, J, 3, 8, hd)
            q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]      # (S,F,J,8,hd)
            if temporal:
                q, k, v = (z.permute(0, 2, 3, 1, 4) for z in (q, k, v))      # (S,J,8,F,hd)
            else:
                q, k, v = (z.permute(0, 1, 3, 2, 4) for z in (q, k, v))      # (S,F,8,J,hd)
            a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v
            a = a.permute(0, 3, 1, 2, 4) if temporal else a.permute(0, 1, 3, 2, 4)   # -> (S,F,J,8,hd)
            ref = a.reshape(S * F * J, C)
            err = (out.double() - ref).abs().max().item()
            print(f"attention simt={simt} S={S} J={J} C={C} temporal={temporal}: max abs err {err:.3e}", flush=True)
            assert err < 2e-5


@rung
def attention():
    _attention(True, [(24, 384), (68, 224), (42, 256)])


@rung
def attention_tc_hands():
    _attention(False, [(42, 256)])


@rung
def attention_tc_face():
    _attention(False, [(68, 224)])


@rung
def attention_tc_body():
    _attention(False, [(24, 384)])


@rung
def attention_tc_big():
    _attention(False, [(24, 384), (68, 224), (42, 256)], S=37)


@rung
def qkv_attention():
    import torch
    torch.manual_seed(0)
    c = _ctx()
    for (J, C) in [(42, 256), (68, 224), (24, 384)]:
        for temporal in (False, True):
            S, F, hd = 7, 27, C // 8
            M = S * F * J
            x = torch.randn(M, C, device="cuda")
            w = (torch.rand(3 * C, C, device="cuda") * 2 - 1) / C ** 0.5 * 2
            b = torch.randn(3 * C, device="cuda") * 0.1
            out = c.qkv_attention(x, w, b, S, J, C, temporal)
            torch.cuda.synchronize()
            t = (x.double() @ w.double().t() + b.double()).reshape(S, F, J, 3, 8, hd)
            q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
            perm = (0, 2, 3, 1, 4) if temporal else (0, 1, 3, 2, 4)
            q, k, v = (z.permute(*perm) for z in (q, k, v))
            a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v
            a = a.permute(0, 3, 1, 2, 4) if temporal else a.permute(0, 1, 3, 2, 4)
            err = (out.double() - a.reshape(M, C)).abs().max().item()
            print(f"qkv_attention J={J} C={C} temporal={temporal}: max abs err {err:.3e}", flush=True)
            assert err < 5e-5


@rung
def model_small():
    import numpy as np
    import torch
    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    for name in ("tiny_B1_H3_K2_d2", "noflip_B2_H1_K2_d2", "cfg1_B2_H1_K1", "small_B2_H2_K3"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        B, H, K, depth, flip = [int(v) for v in g["meta"]]
        sk = H3WBSkeleton()
        args = synthetic.default_args(depth=depth, test_time_augmentation=bool(flip))
        for simt in (True, False):
            m = pafuse_b200.D3DP(args, sk.joints_left, sk.joints_right, sk, is_train=False, num_proposals=H,
                                 sampling_timesteps=K)
            m.load_state_dict(synthetic.synthetic_state_dict(seed=1, depth=depth), strict=False)
            m = m.cuda().eval()
            x2d, x2df = synthetic.synthetic_inputs(B, seed=1)
            noises = synthetic.synthetic_noise(B, H, K, seed=1)
            m.noise_source = lambda k, shape, device: noises[k].to(device)
            m.native_context().set_debug_simt_gemm(simt)
            out = m(x2d.cuda(), None, input_2d_flip=x2df.cuda() if flip else None)
            torch.cuda.synchronize()
            ref = torch.from_numpy(g["out"]).cuda()
            d = (out - ref)
            mpjpe = d.norm(dim=-1).mean().item()
            rel = (d.abs() / ref.abs().clamp_min(1e-2))
            print(f"{name} simt={simt}: max abs {d.abs().max().item():.3e} mpjpe-delta {mpjpe * 1e3:.5f} mm "
                  f"max rel(floor 1e-2) {rel.max().item():.3e} frac>1e-3 {(rel > 1e-3).float().mean().item():.2e}")


def main():
    names = sys.argv[1:] or list(RUNGS)
    if len(names) == 1 and names[0].startswith("--run="):
        RUNGS[names[0][6:]]()
        return
    ok = True
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, f"--run={n}"], timeout=240, capture_output=True, text=True)
            status = "ok" if r.returncode == 0 else f"FAIL rc={r.returncode}"
            out = (r.stdout + r.stderr).strip()
        except subprocess.TimeoutExpired as e:
            status, out = "TIMEOUT", ((e.stdout or b"").decode() + (e.stderr or b"").decode())
        ok &= status == "ok"
        print(f"=== {n}: {status} ({time.time() - t0:.1f}s)\n{out[-4000:]}\n", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
