"""ORACLE tooling — dumps the model / optimizer configuration of the BASELINE.json configs as the
reference's OWN config loader resolves them (utils/config.py: `!include`, `_default` merge, dotted
overrides, ${ENV}), plus the state_dict key -> shape map of the model the LIVE reference builds
from each (common.py:127-130).  Run in the build container only:

    python oracle/make_model_cfgs.py

Output: tests/golden/model_cfgs.json (travels to the GPU box; /root/reference does not).
`utils.config` parses sys.argv at import (utils/config.py:216), hence one subprocess per yml.
"""
import json
import os
import subprocess
import sys

REF = os.environ.get("YAMB_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "model_cfgs.json")

CONFIGS = {
    "mobilenet_v2": "apps/mobilenet/mobilenet_v2_mnas.yml",
    "proxyless_mobile": "apps/mobilenet/proxyless_mobile_mnas.yml",
    "atomnas_c+": "apps/searched/atomnas_c/atomnas_c+.yml",
    "autonl_l": "apps/searched/autonl/autonl_l.yml",
}

CHILD = r"""
import importlib, json, sys, logging
logging.disable(logging.CRITICAL)
from utils.config import FLAGS
import models.mobilenet_base as mb

def plain(v):
    if isinstance(v, dict):
        return {k: plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    return v

lib = importlib.import_module(FLAGS.model)
kw = plain(dict(FLAGS.model_kwparams))
model = lib.Model(**FLAGS.model_kwparams, input_size=FLAGS.image_size)
keys = [k for k in ('model', 'image_size', 'per_gpu_batch_size', 'optimizer', 'alpha', 'momentum',
                    'epsilon', 'eps_inside_sqrt', 'base_lr', 'base_total_batch', 'weight_decay',
                    'weight_decay_method', 'label_smoothing', 'moving_average_decay',
                    'moving_average_decay_adjust', 'moving_average_decay_base_batch',
                    'reset_param_method', 'random_seed', 'bn_calibration', 'bn_calibration_steps',
                    'nl_norm') if k in FLAGS]
out = {'flags': {k: plain(FLAGS[k]) for k in keys}, 'model_kwparams': kw,
       'state_shapes': {k: list(v.shape) for k, v in model.state_dict().items()},
       'n_params': sum(p.numel() for p in model.parameters())}
print('@@' + json.dumps(out))
"""


def main():
    env = dict(os.environ, ARNOLD_OUTPUT="/tmp/yamb_out", DATA_LMDB="/tmp/yamb_lmdb",
               PYTHONPATH=REF)
    res = {}
    for name, yml in CONFIGS.items():
        p = subprocess.run([sys.executable, "-c", CHILD, "app:" + os.path.join(REF, yml)],
                           cwd=REF, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("%s: %s" % (yml, p.stderr[-2000:]))
        line = [l for l in p.stdout.splitlines() if l.startswith("@@")][-1]
        res[name] = json.loads(line[2:])
        res[name]["yml"] = yml
        print(name, res[name]["n_params"], "params,", len(res[name]["state_shapes"]), "state keys")
    with open(OUT, "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
