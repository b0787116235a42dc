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
