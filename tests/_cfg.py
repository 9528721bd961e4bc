"""Shared helper of the config tests: build a model of this package from the kwargs the
reference's own yml loader resolved (tests/golden/model_cfgs.json, written by
oracle/make_model_cfgs.py from apps/**/*.yml)."""
import importlib
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULE_MAP = {"models.mobilenet_supernet": "yet_another_mobilenet_series_b200.mobilenet_supernet",
              "models.searched_network": "yet_another_mobilenet_series_b200.searched_network"}


def load_cfgs():
    with open(os.path.join(ROOT, "tests", "golden", "model_cfgs.json")) as f:
        return json.load(f)


def build_from_cfg(name, seed=None, num_classes=None):
    """What the reference's common.get_model does (common.py:127-141) with `model:` pointed at this
    package: Model(**model_kwparams, input_size=image_size) + init_weights_mnas."""
    from yet_another_mobilenet_series_b200 import mobilenet_base as mb
    cfg = load_cfgs()[name]
    lib = importlib.import_module(MODULE_MAP[cfg["flags"]["model"]])
    kw = dict(cfg["model_kwparams"])
    if num_classes is not None:
        kw["num_classes"] = num_classes
    torch.manual_seed(cfg["flags"]["random_seed"] if seed is None else seed)
    model = lib.Model(**kw, input_size=cfg["flags"]["image_size"])
    if cfg["flags"].get("reset_param_method") == "mnas":
        model.apply(mb.init_weights_mnas)
    return model, cfg
