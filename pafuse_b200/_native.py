"""ctypes binding of the C-ABI library (``include/pafuse_b200.h``).

PyTorch is used only for device memory and streams: every call passes raw
``data_ptr()`` values and the current CUDA stream handle.  There is no CPU or
PyTorch fallback: if the shared library is missing or no sm_100 device is
present the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

import torch

from . import build as _build

MAX_PARTS = 4


class PafuseConfig(ctypes.Structure):
    _fields_ = [
        ("frames", c_int32), ("num_kps", c_int32), ("depth", c_int32), ("heads", c_int32), ("num_parts", c_int32),
        ("part_channels", c_int32 * MAX_PARTS), ("part_num_joints", c_int32 * MAX_PARTS),
        ("part_joints", POINTER(c_int32) * MAX_PARTS), ("flip_perm", POINTER(c_int32)),
        ("scale", c_float), ("max_seqs", c_int32),
    ]


_SIGNATURES = {
    "pafuse_last_error": (c_char_p, []),
    "pafuse_version": (c_char_p, []),
    "pafuse_launch_count": (c_int64, []),
    "pafuse_create": (c_int32, [POINTER(PafuseConfig), POINTER(c_void_p)]),
    "pafuse_destroy": (None, [c_void_p]),
    "pafuse_set_weight": (c_int32, [c_void_p, c_int32, c_char_p, c_void_p, c_int64, c_int32]),
    "pafuse_commit_weights": (c_int32, [c_void_p, c_void_p]),
    "pafuse_pred_parts": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "pafuse_ddim_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_int32, c_int32, c_int32, c_int32, c_double, c_double, c_double, c_float, c_float,
                                   c_float, c_void_p]),
    "pafuse_wb_pose_from_parts": (c_int32, [c_void_p, c_void_p, c_void_p, POINTER(c_int32), c_int64, c_int32, c_void_p]),
    "pafuse_project_to_2d": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "pafuse_aggregate": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pafuse_mpjpe_metrics": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                       c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pafuse_mpjpe_metrics_parts": (c_int32, [c_void_p, c_void_p, c_void_p, POINTER(c_int32), POINTER(c_int32), c_int32,
                                             c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "pafuse_randn": (c_int32, [c_void_p, c_uint64, c_uint64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "pafuse_set_graph_max_seqs": (c_int32, [c_void_p, c_int32]),
    "pafuse_graph_replays": (c_int64, [c_void_p]),
    "pafuse_prepare_clips": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "pafuse_stitch_clips": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int64, c_void_p, c_void_p]),
    "pafuse_keypoints_from_detections": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "pafuse_profile_enable": (c_int32, [c_void_p, c_int32]),
    "pafuse_profile_read": (c_int32, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int64), c_int32]),
    "pafuse_profile_read_bytes": (c_int32, [c_void_p, POINTER(c_double), c_int32]),
    "pafuse_linear": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                c_int32, c_void_p]),
    "pafuse_attention": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pafuse_qkv_attention": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p]),
    "pafuse_set_debug_simt_gemm": (c_int32, [c_void_p, c_int32]),
    "pafuse_set_gemm_cta_group": (c_int32, [c_int32]),
    "pafuse_set_gemm_weight_stationary": (c_int32, [c_int32]),
    "pafuse_mlp_block": (c_int32, [c_void_p] * 12 + [c_int64, c_int32, c_int32, c_void_p]),
    "pafuse_set_fuse_layernorm": (c_int32, [c_void_p, c_int32]),
    "pafuse_set_part_streams": (c_int32, [c_void_p, c_int32, c_void_p]),
    "pafuse_set_fuse_mlp": (c_int32, [c_void_p, c_int32]),
    "pafuse_set_debug_simt_attention": (c_int32, [c_void_p, c_int32]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib_path() -> str:
    """The in-tree library; PAFUSE_LIB names another build of the same sources for A/B measurements."""
    return os.environ.get("PAFUSE_LIB") or _build.LIB_PATH


def load_library(build_if_missing: bool = True):
    """dlopen the in-tree library (building it first when sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and path == _build.LIB_PATH and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on the box and a stale/missing .so
            if not os.path.isfile(path):
                raise RuntimeError(f"pafuse_b200: native library missing and cannot be built: {e}") from e
    if not os.path.isfile(path):
        raise RuntimeError(f"pafuse_b200: native library not found at {path}; run `python -m pafuse_b200.build`")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PafuseError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().pafuse_last_error().decode(errors="replace")
        raise PafuseError(f"{what} failed (rc={rc}): {msg}")


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, device=None):
    t = t.detach()
    if device is not None and t.device != device:
        t = t.to(device)
    return t.to(torch.float32).contiguous()


class NativeContext:
    """Owns one ``pafuse_ctx`` bound to one CUDA device."""

    def __init__(self, frames, num_kps, depth, heads, part_channels, part_joints, flip_perm, scale, max_seqs, device):
        if not torch.cuda.is_available():
            raise PafuseError("pafuse_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = torch.device(device)
        self.num_kps = num_kps
        self.frames = frames
        self.part_channels = list(part_channels)
        cfg = PafuseConfig()
        cfg.frames, cfg.num_kps, cfg.depth, cfg.heads, cfg.num_parts = frames, num_kps, depth, heads, len(part_channels)
        self._keep = []
        for i, (C, joints) in enumerate(zip(part_channels, part_joints)):
            arr = (c_int32 * len(joints))(*joints)
            self._keep.append(arr)
            cfg.part_channels[i] = C
            cfg.part_num_joints[i] = len(joints)
            cfg.part_joints[i] = ctypes.cast(arr, POINTER(c_int32))
        fp = (c_int32 * num_kps)(*flip_perm)
        self._keep.append(fp)
        cfg.flip_perm = ctypes.cast(fp, POINTER(c_int32))
        cfg.scale = float(scale)
        cfg.max_seqs = int(max_seqs)
        handle = c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_create(ctypes.byref(cfg), ctypes.byref(handle)), "pafuse_create")
        self.handle = handle

    def close(self):
        if getattr(self, "handle", None):
            with torch.cuda.device(self.device):
                self.lib.pafuse_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def set_weight(self, part: int, name: str, tensor):
        t = _f32c(tensor)
        on_dev = 1 if t.is_cuda else 0
        if t.is_cuda and t.device != self.device:
            t = t.to(self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_set_weight(self.handle, part, name.encode(), _ptr(t), t.numel(), on_dev),
                  f"pafuse_set_weight({name})")

    def commit_weights(self):
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_commit_weights(self.handle, _stream()), "pafuse_commit_weights")

    # ---- compute
    def pred_parts(self, x2d, x3d, sinus, out=None):
        B, H = x3d.shape[0], x3d.shape[1]
        x2d, x3d, sinus = _f32c(x2d, self.device), _f32c(x3d, self.device), _f32c(sinus, self.device)
        if out is None:
            out = torch.empty_like(x3d)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_pred_parts(self.handle, _ptr(x2d), _ptr(x3d), _ptr(sinus), _ptr(out), B, H, _stream()),
                  "pafuse_pred_parts")
        return out

    def ddim_step(self, x2d, x2d_flip, sinus, img, noise, x0_out, x0_batch_stride, B, H, flip, last,
                  sqrt_recip, sqrt_recipm1, c64, sqrt_an, c, sigma):
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_ddim_step(self.handle, _ptr(x2d), _ptr(x2d_flip), _ptr(sinus), _ptr(img), _ptr(noise),
                                            c_void_p(x0_out), x0_batch_stride, B, H, int(flip), int(last),
                                            sqrt_recip, sqrt_recipm1, c64, sqrt_an, c, sigma, _stream()),
                  "pafuse_ddim_step")

    def wb_pose_from_parts(self, pose, conn_of_joint, mutate_input=True):
        assert pose.is_cuda and pose.dtype == torch.float32 and pose.is_contiguous()
        out = torch.empty_like(pose)
        poses = pose.numel() // (self.num_kps * 3)
        arr = (c_int32 * self.num_kps)(*conn_of_joint)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_wb_pose_from_parts(self.handle, _ptr(pose), _ptr(out), arr, poses,
                                                     1 if mutate_input else 0, _stream()), "pafuse_wb_pose_from_parts")
        return out

    def project_to_2d(self, X, cam):
        X, cam = _f32c(X, self.device), _f32c(cam, self.device)
        n_cams = X.shape[0]
        pts = X.numel() // (3 * n_cams)
        out = torch.empty(tuple(X.shape[:-1]) + (2,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_project_to_2d(self.handle, _ptr(X), _ptr(cam), _ptr(out), n_cams, pts, _stream()),
                  "pafuse_project_to_2d")
        return out

    def aggregate(self, pred, traj, cam, x2d, want_select=True, want_reproj=False):
        pred, x2d, cam = _f32c(pred, self.device), _f32c(x2d, self.device), _f32c(cam, self.device)
        traj = None if traj is None else _f32c(traj, self.device)
        B, K, H, F, J, _ = pred.shape
        jagg = torch.empty((B, K, F, J, 3), dtype=torch.float32, device=self.device)
        pagg = torch.empty_like(jagg)
        sel = torch.empty((B, K, F, J), dtype=torch.int32, device=self.device) if want_select else None
        rep = torch.empty((B, K, H, F, J, 2), dtype=torch.float32, device=self.device) if want_reproj else None
        cam_per_clip = 1 if (cam.dim() == 2 and cam.shape[0] == B and B > 1) else 0
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_aggregate(self.handle, _ptr(pred), _ptr(traj), _ptr(cam), cam_per_clip, _ptr(x2d),
                                            _ptr(jagg), _ptr(pagg), _ptr(sel), _ptr(rep), B, K, H, _stream()),
                  "pafuse_aggregate")
        return jagg, pagg, sel, rep

    # ---- unit-level
    def linear(self, x, w, b, epilogue=0, use_simt=False, y=None):
        x, w, b = _f32c(x, self.device), _f32c(w, self.device), _f32c(b, self.device)
        M, K = x.shape
        N = w.shape[0]
        if y is None:
            y = torch.empty((M, N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_linear(self.handle, _ptr(x), _ptr(w), _ptr(b), _ptr(y), M, N, K, epilogue,
                                         1 if use_simt else 0, _stream()), "pafuse_linear")
        return y

    def attention(self, qkv, S, J, C, temporal):
        qkv = _f32c(qkv, self.device)
        out = torch.empty((qkv.shape[0], C), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_attention(self.handle, _ptr(qkv), _ptr(out), S, J, C, 1 if temporal else 0, _stream()),
                  "pafuse_attention")
        return out

    def qkv_attention(self, x, w, b, S, J, C, temporal):
        x, w, b = _f32c(x, self.device), _f32c(w, self.device), _f32c(b, self.device)
        out = torch.empty((x.shape[0], C), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_qkv_attention(self.handle, _ptr(x), _ptr(w), _ptr(b), _ptr(out), S, J, C,
                                                1 if temporal else 0, _stream()), "pafuse_qkv_attention")
        return out

    def mlp_block(self, a, w1, b1, w2, b2, x, g1, bb1, g0=None, bb0=None, fused=True):
        """(x', a_out): x' = x + fc2(GELU(fc1(a))) [then LN(.; g0, bb0)], a_out = LN(x'; g1, bb1); see pafuse_mlp_block."""
        a, w1, b1, w2, b2 = (_f32c(t, self.device) for t in (a, w1, b1, w2, b2))
        g1, bb1 = _f32c(g1, self.device), _f32c(bb1, self.device)
        g0 = None if g0 is None else _f32c(g0, self.device)
        bb0 = None if bb0 is None else _f32c(bb0, self.device)
        x = _f32c(x, self.device).clone()
        M, C = a.shape
        a_out = torch.empty_like(a)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_mlp_block(self.handle, _ptr(a), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(x), _ptr(g0),
                                            _ptr(bb0), _ptr(g1), _ptr(bb1), _ptr(a_out), M, C, 1 if fused else 0, _stream()),
                  "pafuse_mlp_block")
        return x, a_out

    def set_debug_simt_gemm(self, enable: bool):
        check(self.lib.pafuse_set_debug_simt_gemm(self.handle, 1 if enable else 0), "pafuse_set_debug_simt_gemm")

    def mpjpe_metrics(self, pred, target, traj, cam, x2d, reproj=None):
        """(K, 3+H) SUMS over (b,f,j): J-Best, P-Agg, J-Agg (0 without a 2D target), per-hypothesis root-centred error."""
        pred, target = _f32c(pred, self.device), _f32c(target, self.device)
        x2d = None if x2d is None else _f32c(x2d, self.device)
        traj = None if traj is None else _f32c(traj, self.device)
        cam = None if cam is None else _f32c(cam, self.device)
        reproj = None if reproj is None else _f32c(reproj, self.device)
        B, K, H, F, J, _ = pred.shape
        sums = torch.empty((K, 3 + H), dtype=torch.float64, device=self.device)
        cam_per_clip = 1 if (cam is not None and cam.dim() == 2 and cam.shape[0] == B and B > 1) else 0
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_mpjpe_metrics(self.handle, _ptr(pred), _ptr(target), _ptr(traj), _ptr(cam), cam_per_clip,
                                                _ptr(x2d), _ptr(reproj), _ptr(sums), B, K, H, _stream()),
                  "pafuse_mpjpe_metrics")
        return sums

    def mpjpe_metrics_parts(self, pred, target, part_of_joint, root_of_joint, n_parts):
        """(K, H+1, n_parts) SUMS of the part-centred errors: rows h < H per hypothesis, row H for the mean pose."""
        pred, target = _f32c(pred, self.device), _f32c(target, self.device)
        B, K, H, F, J, _ = pred.shape
        sums = torch.empty((K, H + 1, n_parts), dtype=torch.float64, device=self.device)
        pj = (c_int32 * J)(*[int(v) for v in part_of_joint])
        rj = (c_int32 * J)(*[int(v) for v in root_of_joint])
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_mpjpe_metrics_parts(self.handle, _ptr(pred), _ptr(target), pj, rj, n_parts, _ptr(sums),
                                                      B, K, H, _stream()), "pafuse_mpjpe_metrics_parts")
        return sums

    def randn(self, seed, draw, base, rows, row_len, row_stride, out=None):
        """Counter-based N(0,1) draws: element i of row r is global element ``base + r*row_stride + i`` of draw ``draw``."""
        if out is None:
            out = torch.empty((rows, row_len), dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == rows * row_len
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_randn(self.handle, int(seed) & 0xFFFFFFFFFFFFFFFF, int(draw), int(base), _ptr(out),
                                        int(rows), int(row_len), int(row_stride), _stream()), "pafuse_randn")
        return out

    def set_graph_max_seqs(self, max_seqs: int):
        check(self.lib.pafuse_set_graph_max_seqs(self.handle, int(max_seqs)), "pafuse_set_graph_max_seqs")

    def graph_replays(self) -> int:
        return int(self.lib.pafuse_graph_replays(self.handle))

    # ---- caller-side preparation
    def prepare_clips(self, seq, want_flip=True):
        seq = _f32c(seq, self.device)
        T = seq.shape[0]
        n = (T + self.frames - 1) // self.frames
        clips = torch.empty((n, self.frames, self.num_kps, 2), dtype=torch.float32, device=self.device)
        flip = torch.empty_like(clips) if want_flip else None
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_prepare_clips(self.handle, _ptr(seq), T, _ptr(clips), _ptr(flip), _stream()),
                  "pafuse_prepare_clips")
        return clips, flip

    def stitch_clips(self, pred, T):
        pred = _f32c(pred, self.device)
        n, K, H = pred.shape[0], pred.shape[1], pred.shape[2]
        out = torch.empty((K, H, T, self.num_kps, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_stitch_clips(self.handle, _ptr(pred), n, K, H, T, _ptr(out), _stream()),
                  "pafuse_stitch_clips")
        return out

    def keypoints_from_detections(self, raw, width, height):
        raw = _f32c(raw, self.device)
        T = raw.shape[0]
        kp = torch.empty((T, self.num_kps, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_keypoints_from_detections(self.handle, _ptr(raw), T, int(width), int(height), _ptr(kp),
                                                            _stream()), "pafuse_keypoints_from_detections")
        return kp

    PROFILE_CATEGORIES = ("gemm", "attention", "layernorm", "embed_head", "ddim", "post")

    def profile_enable(self, enable: bool = True):
        check(self.lib.pafuse_profile_enable(self.handle, 1 if enable else 0), "pafuse_profile_enable")

    def profile_read(self):
        """{category: (device ms, algorithmic work, launches)} since profile_enable(True)."""
        n = len(self.PROFILE_CATEGORIES)
        ms, work, cnt = (c_double * n)(), (c_double * n)(), (c_int64 * n)()
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_profile_read(self.handle, ms, work, cnt, n), "pafuse_profile_read")
        return {name: (ms[i], work[i], int(cnt[i])) for i, name in enumerate(self.PROFILE_CATEGORIES)}

    def profile_read_bytes(self):
        """{category: algorithmic DRAM bytes} of the launches recorded since profile_enable(True)."""
        n = len(self.PROFILE_CATEGORIES)
        b = (c_double * n)()
        check(self.lib.pafuse_profile_read_bytes(self.handle, b, n), "pafuse_profile_read_bytes")
        return {name: b[i] for i, name in enumerate(self.PROFILE_CATEGORIES)}

    def set_debug_simt_attention(self, enable: bool):
        check(self.lib.pafuse_set_debug_simt_attention(self.handle, 1 if enable else 0), "pafuse_set_debug_simt_attention")

    def set_gemm_cta_group(self, cta_group: int):
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_set_gemm_cta_group(int(cta_group)), "pafuse_set_gemm_cta_group")

    def set_fuse_layernorm(self, enable: bool):
        check(self.lib.pafuse_set_fuse_layernorm(self.handle, 1 if enable else 0), "pafuse_set_fuse_layernorm")

    def set_fuse_mlp(self, enable: bool):
        check(self.lib.pafuse_set_fuse_mlp(self.handle, 1 if enable else 0), "pafuse_set_fuse_mlp")

    def set_part_streams(self, enable: bool, shares=None):
        arr = None
        if shares is not None:
            arr = (c_int32 * len(shares))(*[int(v) for v in shares])
        check(self.lib.pafuse_set_part_streams(self.handle, 1 if enable else 0, arr), "pafuse_set_part_streams")

    def set_gemm_weight_stationary(self, enable: bool):
        with torch.cuda.device(self.device):
            check(self.lib.pafuse_set_gemm_weight_stationary(1 if enable else 0), "pafuse_set_gemm_weight_stationary")

    def launch_count(self) -> int:
        return int(self.lib.pafuse_launch_count())
