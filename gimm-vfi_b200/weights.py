"""Seeded random-init weights of the GIMM-VFI-R architecture (there are no
checkpoints and no network here; bench.py and the tests use these).  Scales
follow the reference constructors so the random network is as well conditioned
as a freshly-constructed reference model: PyTorch-default conv init
(uniform ±1/sqrt(fan_in)), kaiming-normal fan_out for the RAFT encoders
(raft/extractor.py:156-163), SIREN init for HypoNet (modules/utils.py:37-44),
PReLU 0.25, BatchNorm statistics randomised so that BN folding is exercised."""
import math

import torch

from .arch import param_spec_f, param_spec_r


def random_state_dict(seed: int = 0) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def U(shape, b):
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    for key, shape, dt in param_spec_r():
        leaf = key.rsplit(".", 1)[-1]
        if key == "g_filter":
            t = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32).div(16).reshape(shape)
        elif key in ("alpha_v", "alpha_fe"):
            t = 0.5 + torch.rand(shape, generator=g)
        elif key.startswith("hyponet."):
            fan_in = shape[0] - 1
            first = key.endswith("wb0")
            b = 1.0 / fan_in if first else math.sqrt(6.0 / fan_in)
            t = U(shape, b)  # weight rows and the bias row share the SIREN bound
        elif leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 4:
            cout, cin, kh, kw = shape
            if key.startswith("flow_estimator.fnet") or key.startswith("flow_estimator.cnet"):
                t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (cout * kh * kw))
            else:
                t = U(shape, 1.0 / math.sqrt(cin * kh * kw))
        elif leaf == "bias":
            parent = key[: -len(".bias")]
            wkey = parent + ".weight"
            wshape = next((s for k, s, _ in param_spec_r() if k == wkey), None)
            if wshape is not None and len(wshape) == 4:
                t = U(shape, 1.0 / math.sqrt(wshape[1] * wshape[2] * wshape[3]))
            else:  # BatchNorm bias
                t = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 1:
            parent = key[: -len(".weight")]
            is_bn = any(k == parent + ".running_mean" for k, _, _ in param_spec_r())
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if is_bn else (0.25 + 0.05 * torch.randn(shape, generator=g))
        else:
            raise KeyError(key)
        sd[key] = t.to(getattr(torch, dt)).contiguous()
    # norm3 and downsample.1 are one module object in the reference (raft/extractor.py:44-47)
    for key in list(sd):
        if ".norm3." in key:
            sd[key.replace(".norm3.", ".downsample.1.")] = sd[key]
    return sd


def _flow_estimator_init(key: str, shape, dt: str, spec: dict, g: torch.Generator) -> torch.Tensor:
    """Seeded init of one `flow_estimator.*` (FlowFormer) tensor of GIMM-VFI-F, shaped like a trained network's so that fixtures exercise
    every branch (torch-default uniform fan-in bounds for Linear / Conv weights and biases, LayerNorm scales near 1, GMA gamma != 0)."""
    U = lambda sh, b: (torch.rand(sh, generator=g) * 2 - 1) * b
    leaf = key.rsplit(".", 1)[-1]
    if dt == "int64":
        return torch.zeros(shape, dtype=torch.int64)      # att.pos_emb.rel_ind: the positional term is disabled in the reference forward (gma.py:64-70)
    if leaf == "latent_tokens":
        return torch.randn(shape, generator=g)
    if leaf == "gamma":
        return torch.full(shape, 0.35)
    is_norm = ".norm" in key.rsplit(".", 1)[0].rsplit(".", 1)[-1] or key.rsplit(".", 2)[-2].startswith("norm")
    if len(shape) == 1 and is_norm:
        return (1.0 + 0.1 * torch.randn(shape, generator=g)) if leaf == "weight" else 0.05 * torch.randn(shape, generator=g)
    if leaf == "weight" and len(shape) >= 2:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return U(shape, 1.0 / math.sqrt(fan_in))
    if leaf == "bias":
        wshape = spec.get(key[: -len("bias")] + "weight")
        fan_in = 1
        for d in (wshape[1:] if wshape else shape):
            fan_in *= d
        return U(shape, 1.0 / math.sqrt(fan_in))
    raise KeyError(key)


def random_state_dict_f(seed: int = 0, flow_estimator: dict = None) -> dict:
    """Seeded GIMM-VFI-F state_dict: the decoder / GIMM half takes the SAME tensors as random_state_dict(seed) (identical layout,
    arch.param_spec_f); `flow_estimator.*` (FlowFormer) comes from the caller (e.g. a reference module's own init) or from the seeded
    init above — reproducible anywhere, so GPU tests rebuild exactly the weights the reference ran when the fixtures were made."""
    r = random_state_dict(seed)
    g = torch.Generator().manual_seed(seed + 1000)
    spec = {k: s for k, s, _ in param_spec_f()}
    sd = {}
    for key, shape, dt in param_spec_f():
        if key.startswith("flow_estimator."):
            t = flow_estimator[key].detach().clone() if flow_estimator is not None else _flow_estimator_init(key, shape, dt, spec, g)
            assert tuple(t.shape) == tuple(shape), key
            sd[key] = t
        else:
            sd[key] = r[key]
    return sd
