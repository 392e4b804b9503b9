"""Config and checkpoint compatibility.

The reference reads a hydra ``DictConfig`` (``@hydra.main(config_path="config", config_name="config")``,
``main_h3wb.py:567``) by attribute and loads ``pafuse_model.bin`` written by ``common/logging.py:83-115``.
hydra / omegaconf are optional here: ``load_config`` parses the same ``config/config.yaml`` layout (two-level YAML,
``group.key`` overrides like hydra's command line) into attribute-style namespaces, filling the keys the denoising
path reads (``diffusionpose.py:62-103,140-153``; ``main_h3wb.py:687-688``) with the file's shipped defaults when absent.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

__all__ = ["load_config", "model_from_config", "load_checkpoint", "HOT_PATH_DEFAULTS"]

# keys of config/config.yaml the inference path reads, with the values the reference ships
HOT_PATH_DEFAULTS = {
    "general": {"part_based_model": True, "evaluate": "best_epoch.bin"},
    "data": {"dataset": "h3wb", "num_kps": 134, "merge_hands": True},
    "model": {"diff_model": "MixSTE2", "number_of_frames": 27, "batch_size": 1024, "test_time_augmentation": True,
              "cs": 288, "dep": 8, "input_size": 5},
    "ft2d": {"scale": 1.0, "timestep": 1000, "sampling_timesteps": 5, "num_proposals": 10, "debug": False, "p2": False},
    "in_the_wild": {"video_path": ""},
}


def _coerce(text: str):
    low = text.lower()
    if low in ("true", "false"):
        return low == "true"
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    return text


def load_config(path=None, overrides=()):
    """YAML file (optional) + ``group.key=value`` overrides -> nested ``SimpleNamespace`` with attribute access."""
    cfg = {g: dict(kv) for g, kv in HOT_PATH_DEFAULTS.items()}
    if path:
        import yaml
        with open(path, "r") as f:
            data = yaml.safe_load(f) or {}
        for group, kv in data.items():
            if isinstance(kv, dict):
                cfg.setdefault(group, {}).update(kv)
            else:
                cfg[group] = kv
    for ov in overrides:
        key, _, val = ov.partition("=")
        group, _, name = key.partition(".")
        if not name:
            raise ValueError(f"override '{ov}' must look like group.key=value")
        cfg.setdefault(group, {})[name] = _coerce(val)
    return SimpleNamespace(**{g: SimpleNamespace(**kv) if isinstance(kv, dict) else kv for g, kv in cfg.items()})


def load_checkpoint(path, map_location="cpu"):
    """``pafuse_model.bin`` (``logging.py:94-104``: dict with 'model_pos', 'epoch', 'lr', 'optimizer'[, 'random_state'])
    or a bare state_dict -> the model state_dict (``module.`` prefixes are handled by ``D3DP.load_state_dict``)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    return ckpt["model_pos"] if isinstance(ckpt, dict) and "model_pos" in ckpt else ckpt


def model_from_config(args, dataset, checkpoint=None, device="cuda"):
    """``main_h3wb.py:687-714`` without DataParallel: build the eval model from the config and load a checkpoint."""
    from .diffusionpose import D3DP
    sym = dataset.keypoints_metadata["keypoints_symmetry"]
    model = D3DP(args, list(sym[0]), list(sym[1]), dataset, is_train=False, num_proposals=args.ft2d.num_proposals,
                 sampling_timesteps=args.ft2d.sampling_timesteps)
    if checkpoint is not None:
        sd = load_checkpoint(checkpoint) if isinstance(checkpoint, str) else checkpoint
        model.load_state_dict(sd, strict=False)
    return model.to(device).eval()
