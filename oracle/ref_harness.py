"""Import the UNMODIFIED reference from /root/reference for pinning the oracle.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container); the GPU box never has it, so nothing reachable from ``-m gpu`` tests,
``smoke()`` or ``bench.py`` calls into this file.  Used by
``tests/golden/make_golden.py`` (writes the committed fixtures) and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).

Shims: a fake ``timm`` (only ``DropPath`` must exist and it is
never instantiated at inference), attribute-style args, a skeleton stand-in for
the dataset object, ``device='cpu'`` and a no-op ``Tensor.cuda`` while the
reference runs (it hard-codes ``.cuda()`` at diffusionpose.py:288).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PAFUSE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "common", "diffusionpose.py"))


def _install_timm_stub():
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_pafuse_stub", False):
        return
    timm = types.ModuleType("timm")
    timm._pafuse_stub = True
    data = types.ModuleType("timm.data")
    data.IMAGENET_DEFAULT_MEAN = data.IMAGENET_DEFAULT_STD = (0.5, 0.5, 0.5)
    models = types.ModuleType("timm.models")
    helpers = types.ModuleType("timm.models.helpers")
    helpers.load_pretrained = lambda *a, **k: None
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):  # identity at inference; never built when drop_path == 0
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: (x, x)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    registry = types.ModuleType("timm.models.registry")
    registry.register_model = lambda f: f
    for name, mod in (("timm", timm), ("timm.data", data), ("timm.models", models), ("timm.models.helpers", helpers),
                      ("timm.models.layers", layers), ("timm.models.registry", registry)):
        sys.modules[name] = mod


def import_reference():
    """Returns the reference modules (diffusionpose, mixste, utils, camera, loss)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_timm_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    mods = {}
    for name in ("diffusionpose", "mixste", "utils", "camera", "loss"):
        mods[name] = importlib.import_module(f"common.{name}")
    return types.SimpleNamespace(**mods)


@contextlib.contextmanager
def cpu_cuda_shim():
    """Neutralise Tensor.cuda() while the reference runs on the CPU."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


@contextlib.contextmanager
def injected_noise(noises):
    """Make torch.randn / randn_like return the given tensors in order, so the
    reference sampler consumes exactly the injected draws (diffusionpose.py:283,308)."""
    it = iter(noises)
    o_randn, o_like = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    try:
        yield
    finally:
        torch.randn, torch.randn_like = o_randn, o_like


def build_reference_model(args, skeleton, state_dict, num_proposals, sampling_timesteps):
    ref = import_reference()
    model = ref.diffusionpose.D3DP(args, skeleton.joints_left, skeleton.joints_right, skeleton, is_train=False,
                                   num_proposals=num_proposals, sampling_timesteps=sampling_timesteps)
    model.device = "cpu"
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all(not m.startswith("pose_estimator") for m in missing), missing
    model.eval()
    return model, ref


def reference_forward(model, x2d, x2d_flip, noises):
    with torch.no_grad(), cpu_cuda_shim(), injected_noise(noises):
        return model(x2d, None, input_2d_flip=x2d_flip)
