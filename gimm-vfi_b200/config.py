"""Config loading that mirrors the reference's surface (`config.arch` node consumed by
create_model, src/models/__init__.py:15-37; yaml layout configs/gimmvfi/*.yaml) without
depending on omegaconf."""
import copy

import yaml


class ConfigNode(dict):
    """dict with attribute access and .copy(), enough of an OmegaConf node for this path."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return ConfigNode({k: (v.copy() if isinstance(v, ConfigNode) else copy.deepcopy(v)) for k, v in self.items()})


def _wrap(x):
    if isinstance(x, dict):
        return ConfigNode({k: _wrap(v) for k, v in x.items()})
    return x


_ARCH_DEFAULTS = {  # GIMMVFIConfig / HypoNetConfig dataclass defaults (configs.py:39-57, module_config.py:28-41)
    "type": "gimmvfi_r", "ema": None, "ema_value": None, "fwarp_type": "linear", "rec_weight": 0.1, "raft_iter": 20,
    "coord_range": [-1.0, 1.0], "modulated_layer_idxs": None,
    "hyponet": {"type": "mlp", "n_layer": 5, "hidden_dim": [128], "use_bias": True, "input_dim": 3, "output_dim": 2,
                "output_bias": 0.5, "normalize_weight": True, "linear_interpo": False,
                "activation": {"type": "siren", "siren_w0": 1.0},
                "initialization": {"weight_init_type": "siren", "bias_init_type": "siren"}},
}


def _merge(base, over):
    out = dict(base)
    for k, v in (over or {}).items():
        out[k] = _merge(base[k], v) if isinstance(v, dict) and isinstance(base.get(k), dict) else v
    return out


def default_arch_config() -> ConfigNode:
    return _wrap(copy.deepcopy(_ARCH_DEFAULTS))


def load_config(path: str) -> ConfigNode:
    """yaml -> config with `.arch` merged over the dataclass defaults."""
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    raw["arch"] = _merge(_ARCH_DEFAULTS, raw.get("arch"))
    return _wrap(raw)
