"""CPU: the four BASELINE.json model configs, as the reference's own config loader resolves its
ymls (fixture tests/golden/model_cfgs.json from oracle/make_model_cfgs.py), build through this
package's plugin modules into models whose state_dict keys AND shapes equal those of the model the
live reference built from the same yml (common.py:127-130) — the checkpoint compatibility contract
of SURVEY.md §8b."""
import pytest

from _cfg import build_from_cfg, load_cfgs

NAMES = ["mobilenet_v2", "proxyless_mobile", "atomnas_c+", "autonl_l"]


@pytest.mark.parametrize("name", NAMES)
def test_state_dict_matches_live_reference(name):
    model, cfg = build_from_cfg(name)
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    want = cfg["state_shapes"]
    assert sorted(got) == sorted(want)      # same key set (the fixture is stored key-sorted)
    assert got == want
    assert sum(p.numel() for p in model.parameters()) == cfg["n_params"]


def test_fixture_covers_baseline_configs():
    cfgs = load_cfgs()
    assert set(NAMES) <= set(cfgs)
    assert cfgs["mobilenet_v2"]["model_kwparams"]["active_fn"] == "nn.ReLU"
    assert cfgs["atomnas_c+"]["model_kwparams"]["block"] == "InvertedResidualChannelsFused"
    assert cfgs["atomnas_c+"]["model_kwparams"]["se_ratio"] == 0.5
    assert cfgs["autonl_l"]["model_kwparams"]["active_fn"] == "nn.Swish"
