"""TEST INFRASTRUCTURE — not product code.

Imports the *unmodified* reference GIMM-VFI-R modules from /root/reference by
path so they can run on CPU in the build container, and exposes
``build_reference_model()``.  Only usable where /root/reference exists (the
build container); nothing on the GPU box may import this file.

Shims:
  1. ``models`` / ``models.generalizable_INR`` are registered as bare namespace
     packages so the reference ``__init__`` files (which trip a Python>=3.11
     dataclass mutable-default error at configs.py:24,45 /
     modules/module_config.py:37-38) are never executed; ``configs`` and
     ``modules.module_config`` are replaced by bare classes.
  2. ``cupy`` is stubbed (softsplat.py:12 imports it, :263 uses ``@cupy.memoize``).
  3. ``omegaconf`` is stubbed (hyponet.py:16,37 uses ``OmegaConf.to_object``).
  4. ``softsplat_func.forward`` is ``assert False`` on CPU (softsplat.py:439-440);
     it is replaced by a scatter-add restatement of the CUDA kernel at
     softsplat.py:376-421.
  5. ``initialize_RAFT`` hard-loads pretrained_ckpt/raft-things.pth
     (raft/__init__.py:15); replaced by plain construction.
Everything else (raft/*, modules/{fi_components,fi_utils,hyponet,coord_sampler},
gimmvfi_r.py) is the reference's own code, imported unmodified.
"""
import argparse
import importlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("GIMMVFI_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REF_ROOT, "src")
_PKG = os.path.join(_SRC, "models", "generalizable_INR")


def reference_available() -> bool:
    return os.path.isdir(_PKG)


class AttrDict(dict):
    """Minimal OmegaConf-node stand-in: attribute access + .copy()."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return AttrDict({k: (v.copy() if isinstance(v, AttrDict) else v) for k, v in self.items()})


def default_arch_config() -> AttrDict:
    """configs/gimmvfi/gimmvfi_r_arb.yaml:7-27 merged with the dataclass defaults
    (configs.py:39-57, modules/module_config.py:28-41)."""
    return AttrDict(
        type="gimmvfi_r",
        ema=None,
        ema_value=None,
        fwarp_type="linear",
        rec_weight=0.1,
        raft_iter=20,
        coord_range=[-1.0, 1.0],
        modulated_layer_idxs=[1],
        hyponet=AttrDict(
            type="mlp",
            n_layer=5,
            hidden_dim=[128],
            use_bias=True,
            input_dim=3,
            output_dim=2,
            output_bias=0.5,
            activation=AttrDict(type="siren", siren_w0=1.0),
            initialization=AttrDict(weight_init_type="siren", bias_init_type="siren"),
            normalize_weight=True,
            linear_interpo=False,
        ),
    )


def _splat_forward_cpu(tenIn, tenFlow):
    """Restatement of the ``softsplat_out`` kernel (modules/softsplat.py:376-421)
    with index_put_(accumulate=True) instead of atomicAdd."""
    N, C, H, W = tenIn.shape
    out = tenIn.new_zeros(N, C, H, W)
    gx = torch.arange(W, dtype=tenIn.dtype).view(1, 1, W).expand(N, H, W)
    gy = torch.arange(H, dtype=tenIn.dtype).view(1, H, 1).expand(N, H, W)
    fx = gx + tenFlow[:, 0]
    fy = gy + tenFlow[:, 1]
    finite = torch.isfinite(fx) & torch.isfinite(fy)
    fx = torch.where(finite, fx, torch.zeros_like(fx))
    fy = torch.where(finite, fy, torch.zeros_like(fy))
    x0 = torch.floor(fx).long()
    y0 = torch.floor(fy).long()
    x1 = x0 + 1
    y1 = y0 + 1
    wnw = (x1.to(fx.dtype) - fx) * (y1.to(fy.dtype) - fy)
    wne = (fx - x0.to(fx.dtype)) * (y1.to(fy.dtype) - fy)
    wsw = (x1.to(fx.dtype) - fx) * (fy - y0.to(fy.dtype))
    wse = (fx - x0.to(fx.dtype)) * (fy - y0.to(fy.dtype))
    nidx = torch.arange(N).view(N, 1, 1).expand(N, H, W)
    flat_out = out.permute(0, 2, 3, 1).reshape(N * H * W, C)  # view of a copy? -> use explicit buffer
    buf = torch.zeros(N * H * W, C, dtype=tenIn.dtype)
    src = tenIn.permute(0, 2, 3, 1).reshape(N * H * W, C)
    for xx, yy, ww in ((x0, y0, wnw), (x1, y0, wne), (x0, y1, wsw), (x1, y1, wse)):
        ok = finite & (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        lin = (nidx * H + yy.clamp(0, H - 1)) * W + xx.clamp(0, W - 1)
        okf = ok.reshape(-1)
        buf.index_add_(0, lin.reshape(-1)[okf], src[okf] * ww.reshape(-1, 1)[okf])
    del flat_out
    return buf.view(N, H, W, C).permute(0, 3, 1, 2).contiguous()


_loaded = {}


def load_reference_modules():
    """Returns a dict of the reference's modules (gimmvfi_r, raft, ...)."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    # (2) cupy stub
    if "cupy" not in sys.modules:
        cupy = types.ModuleType("cupy")
        cupy.memoize = lambda **kw: (lambda f: f)
        cupy.ndarray = type("ndarray", (), {})
        cupy.int32 = int
        cupy.float32 = float
        sys.modules["cupy"] = cupy
    # (3) omegaconf stub
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def to_object(x):
                return list(x)

        oc.OmegaConf = OmegaConf
        oc.MISSING = "???"
        sys.modules["omegaconf"] = oc

    # (1) bare namespace packages
    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    root = "gimmvfi_reference"
    ns(root, os.path.join(_SRC, "models"))
    ns(root + ".generalizable_INR", _PKG)
    ns(root + ".generalizable_INR.modules", os.path.join(_PKG, "modules"))
    ns(root + ".generalizable_INR.raft", os.path.join(_PKG, "raft"))
    cfg = types.ModuleType(root + ".generalizable_INR.configs")
    cfg.GIMMVFIConfig = type("GIMMVFIConfig", (), {})
    cfg.GIMMConfig = type("GIMMConfig", (), {})
    cfg.HypoNetConfig = type("HypoNetConfig", (), {})
    sys.modules[cfg.__name__] = cfg
    mcfg = types.ModuleType(root + ".generalizable_INR.modules.module_config")
    mcfg.HypoNetConfig = cfg.HypoNetConfig
    sys.modules[mcfg.__name__] = mcfg

    raft_pkg = sys.modules[root + ".generalizable_INR.raft"]
    raft_mod = importlib.import_module(root + ".generalizable_INR.raft.raft")
    raft_corr = importlib.import_module(root + ".generalizable_INR.raft.corr")

    # (5) RAFT without checkpoint load  (raft/__init__.py:7-24)
    def initialize_RAFT(model_path=None, device="cpu"):
        args = argparse.ArgumentParser()
        args.raft_model = model_path
        args.small = False
        args.mixed_precision = False
        args.alternate_corr = False
        return raft_mod.RAFT(args)

    raft_pkg.initialize_RAFT = initialize_RAFT
    raft_pkg.RAFT = raft_mod.RAFT

    softsplat_mod = importlib.import_module(root + ".generalizable_INR.modules.softsplat")

    # (4) CPU splat
    class _SplatCPU:
        @staticmethod
        def apply(tenIn, tenFlow):
            return _splat_forward_cpu(tenIn.float(), tenFlow.float())

    softsplat_mod.softsplat_func = _SplatCPU

    gimmvfi_r = importlib.import_module(root + ".generalizable_INR.gimmvfi_r")
    fi_utils = importlib.import_module(root + ".generalizable_INR.modules.fi_utils")
    fi_components = importlib.import_module(root + ".generalizable_INR.modules.fi_components")
    hyponet = importlib.import_module(root + ".generalizable_INR.modules.hyponet")
    coord_sampler = importlib.import_module(root + ".generalizable_INR.modules.coord_sampler")
    # fi_utils.warp uses a module-global `device` (fi_utils.py:15) — CPU here.
    fi_utils.device = torch.device("cpu")
    _loaded.update(
        gimmvfi_r=gimmvfi_r,
        raft=raft_mod,
        raft_corr=raft_corr,
        softsplat=softsplat_mod,
        fi_utils=fi_utils,
        fi_components=fi_components,
        hyponet=hyponet,
        coord_sampler=coord_sampler,
    )
    return _loaded


def build_reference_model(state_dict=None, seed=0):
    """Constructs the reference GIMMVFI_R (gimmvfi_r.py:34) on CPU in eval mode."""
    mods = load_reference_modules()
    torch.manual_seed(seed)
    model = mods["gimmvfi_r"].GIMMVFI_R(default_arch_config())
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model


GIMM_KEY_PREFIXES = ("cnn_encoder.", "res_conv.", "hyponet.", "g_filter", "alpha_v", "alpha_fe")


def build_reference_gimm(state_dict=None, seed=0):
    """Constructs the reference GIMM (gimm.py:25, the motion-modelling network alone) on CPU in eval mode.  `state_dict` may be
    a full GIMM-VFI-R state_dict: only gimm.py's keys are taken (same names, gimmvfi_r.py:86-111)."""
    mods = load_reference_modules()
    gimm = importlib.import_module("gimmvfi_reference.generalizable_INR.gimm")
    torch.manual_seed(seed)
    model = gimm.GIMM(default_arch_config())
    if state_dict is not None:
        sub = {k: v for k, v in state_dict.items() if k.startswith(GIMM_KEY_PREFIXES)}
        model.load_state_dict(sub, strict=True)
    model.eval()
    return model
