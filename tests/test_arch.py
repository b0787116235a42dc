import json
import os

import torch

from conftest import GOLDEN_DIR
from gimmvfi_b200.arch import param_spec_r
from gimmvfi_b200.weights import random_state_dict


def test_param_spec_matches_reference_state_dict():
    """The 414-key layout dumped from the reference model."""
    with open(os.path.join(GOLDEN_DIR, "state_dict_spec_r.json")) as f:
        ref = json.load(f)
    mine = param_spec_r()
    assert len(mine) == len(ref) == 414
    for (k, s, d), (k2, s2, d2) in zip(ref, mine):
        assert (k, tuple(s), d) == (k2, tuple(s2), d2)
    import math

    n = sum(math.prod(s) for _, s, _ in mine)  # math.prod(()) == 1 for the int64 scalars
    assert n == 19789980


def test_random_state_dict_deterministic():
    a, b = random_state_dict(7), random_state_dict(7)
    assert list(a) == [k for k, _, _ in param_spec_r()]
    assert all(torch.equal(a[k], b[k]) for k in a)
    k = "flow_estimator.cnet.layer2.0."
    assert torch.equal(a[k + "norm3.weight"], a[k + "downsample.1.weight"])


def test_gimm_state_dict_matches_reference_dump():
    """GIMM standalone (gimm.py:25-80): same 36 keys, order and shapes as the reference module's state_dict"""
    import json, os
    from conftest import GOLDEN_DIR
    from gimmvfi_b200.gimm import GIMM, param_spec_gimm

    spec = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_spec_gimm.json")))
    assert [[k, list(s)] for k, s, _ in param_spec_gimm()] == spec
    m = GIMM()
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(list(sd[k].shape) == s for k, s in spec)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)


def test_param_spec_f_matches_reference_dump():
    """GIMM-VFI-F: the packaged 639-key table == the dump of the reference GIMMVFI_F module (oracle/make_golden_f.py); the
    synthesis half is the GIMM-VFI-R layout without the three feature projections."""
    import json, os
    from conftest import GOLDEN_DIR
    from gimmvfi_b200.arch import R_ONLY_PREFIXES, param_spec_f, param_spec_r

    ref = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_spec_f.json")))
    mine = param_spec_f()
    assert len(mine) == len(ref) == 639
    assert [[k, list(s), d] for k, s, d in mine] == ref
    shared = [k for k, _, _ in param_spec_r() if not k.startswith(R_ONLY_PREFIXES)]
    assert sorted(k for k, _, _ in mine if not k.startswith("flow_estimator.")) == sorted(shared)


def test_yaml_configs_dispatch_to_the_three_boundary_classes():
    """configs/gimmvfi/*.yaml carry the reference's `arch:` blocks; create_model (src/models/__init__.py:15-37) dispatches them to the
    drop-in classes, each exposing the reference's state_dict (414 / 639 tensors)."""
    import os

    from gimmvfi_b200 import create_model, load_config

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "gimmvfi")
    for fn, cls, n in (("gimmvfi_r_arb.yaml", "GIMMVFI_R", 414), ("gimmvfi_f_arb.yaml", "GIMMVFI_F", 639)):
        m, ema = create_model(load_config(os.path.join(root, fn)).arch)
        assert type(m).__name__ == cls and ema is None and len(m.state_dict()) == n
