"""TEST INFRASTRUCTURE — not product code.  The GIMM-VFI-F counterpart of ref_shim.py.

Imports the *unmodified* reference GIMMVFI_F (gimmvfi_f.py) and FlowFormer (flowformer/core/**) from /root/reference so they run
on CPU in the build container.  On top of ref_shim's five shims:

  6. ``timm`` (requirements.txt:74 pins 0.4.12; absent here, no network).  What the reference takes from it:
       * ``timm.models.layers.{Mlp, DropPath, to_2tuple, trunc_normal_}`` and ``timm.models.vision_transformer.Attention``
         (twins.py:22-24, encoder.py:27, cnn.py:4, decoder.py:19) — four trivial utilities, restated below from their published
         definitions (Mlp = fc1 -> GELU -> drop -> fc2 -> drop; DropPath = identity at inference);
       * ``timm.create_model("twins_svt_large", pretrained=...)`` (encoders.py:10) — the Twins-SVT-L architecture.  The reference
         vendors timm's twins.py almost verbatim (LatentCostFormer/twins.py:814-926 LocallyGroupedAttn / GlobalSubSampleAttn,
         :1028-1098 Block, :1100-1150 PosConv / PatchEmbed, :1152-1289 Twins); the stub builds THAT class with timm 0.4.12's
         ``twins_svt_large`` arguments (patch_size 4, embed_dims 128/256/512/1024, heads 4/8/16/32, mlp_ratios 4, depths 2/2/18/2,
         wss 7, sr_ratios 8/4/2/1).  So the Twins arithmetic executed here is the reference tree's own copy of it.
         PARITY UNPINNED AT THE TIMM BOUNDARY: whether timm 0.4.12's file differs from the vendored copy cannot be checked offline.
     ``pretrained=True`` would download ImageNet weights: ignored (random init / the caller's state_dict).
  7. ``yacs.config.CfgNode`` (configs/submission.py:1): an attribute dict with ``clone()``.
  8. ``turtle`` (convnext.py:1 ``from turtle import forward`` — an IDE auto-import artefact; needs tkinter): empty stub.
  9. ``initialize_Flowformer`` hard-loads pretrained_ckpt/flowformer_sintel.pth (flowformer/__init__.py:10): plain construction.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

import ref_shim
from ref_shim import _PKG, AttrDict, default_arch_config  # noqa: F401

_loaded = {}


class CfgNode(dict):
    """yacs.config.CfgNode stand-in: nested attribute access, clone(), keys()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return CfgNode({k: (v.clone() if isinstance(v, CfgNode) else v) for k, v in self.items()})


def _install_stubs():
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        ycfg = types.ModuleType("yacs.config")
        ycfg.CfgNode = CfgNode
        yacs.config = ycfg
        sys.modules["yacs"], sys.modules["yacs.config"] = yacs, ycfg
    if "turtle" not in sys.modules:
        t = types.ModuleType("turtle")
        t.forward = lambda *a, **k: None
        sys.modules["turtle"] = t
    if "timm" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    vit = types.ModuleType("timm.models.vision_transformer")
    registry = types.ModuleType("timm.models.registry")
    helpers = types.ModuleType("timm.models.helpers")
    data = types.ModuleType("timm.data")

    class Mlp(nn.Module):  # timm/models/layers/mlp.py (0.4.12)
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features)
            self.drop = nn.Dropout(drop)

        def forward(self, x):
            return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))

    class DropPath(nn.Module):  # identity at inference
        def __init__(self, drop_prob=None):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training, "stub DropPath is inference-only"
            return x

    class Attention(nn.Module):  # only constructed for ws=None blocks, which Twins-SVT never builds
        def __init__(self, *a, **k):
            super().__init__()
            raise NotImplementedError("timm ViT Attention is not on the GIMM-VFI-F path")

    layers.Mlp, layers.DropPath = Mlp, DropPath
    layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    layers.activations = types.ModuleType("timm.models.layers.activations")
    vit.Attention = Attention
    registry.register_model = lambda f: f
    helpers.build_model_with_cfg = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("timm build_model_with_cfg"))
    helpers.overlay_external_default_cfg = lambda *a, **k: None
    data.IMAGENET_DEFAULT_MEAN, data.IMAGENET_DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    def create_model(name, pretrained=False, **kw):
        assert name == "twins_svt_large", name
        tw = importlib.import_module("gimmvfi_reference.generalizable_INR.flowformer.core.FlowFormer.LatentCostFormer.twins")

        class TimmBlock(nn.Module):
            """timm 0.4.12 twins.Block: the vendored Block (twins.py:1028-1098) was extended with a `context` argument that the
            vendored non-RPE attention classes do not accept, i.e. the reference's encoders run timm's original two-argument Block:
            x + attn(norm1(x), size); x + mlp(norm2(x)) with LocallyGroupedAttn (ws > 1) / GlobalSubSampleAttn (ws == 1)."""

            def __init__(self, dim, num_heads, mlp_ratio=4.0, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU,
                         norm_layer=nn.LayerNorm, sr_ratio=1, ws=None):
                super().__init__()
                self.norm1 = norm_layer(dim)
                assert ws is not None
                self.attn = tw.GlobalSubSampleAttn(dim, num_heads, attn_drop, drop, sr_ratio) if ws == 1 else \
                    tw.LocallyGroupedAttn(dim, num_heads, attn_drop, drop, ws)
                self.drop_path = nn.Identity()
                self.norm2 = norm_layer(dim)
                self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

            def forward(self, x, size):
                x = x + self.drop_path(self.attn(self.norm1(x), size))
                return x + self.drop_path(self.mlp(self.norm2(x)))

        return tw.Twins(block_cls=TimmBlock, patch_size=4, embed_dims=[128, 256, 512, 1024], num_heads=[4, 8, 16, 32], mlp_ratios=[4, 4, 4, 4],
                        depths=[2, 2, 18, 2], wss=[7, 7, 7, 7], sr_ratios=[8, 4, 2, 1])

    timm.create_model = create_model
    timm.models, timm.data = models, data
    models.layers, models.vision_transformer, models.registry, models.helpers = layers, vit, registry, helpers
    for m in (timm, models, layers, vit, registry, helpers, data, layers.activations):
        sys.modules[m.__name__] = m


def load_reference_f():
    if _loaded:
        return _loaded
    ref_shim.load_reference_modules()   # shims 1-5 + the namespace packages
    _install_stubs()
    root = "gimmvfi_reference.generalizable_INR"
    pk = types.ModuleType(root + ".flowformer")
    pk.__path__ = [os.path.join(_PKG, "flowformer")]
    sys.modules[pk.__name__] = pk
    for sub in ("configs", "core", "core.FlowFormer", "core.FlowFormer.LatentCostFormer", "core.utils"):
        m = types.ModuleType(root + ".flowformer." + sub)
        m.__path__ = [os.path.join(_PKG, "flowformer", *sub.split("."))]
        sys.modules[m.__name__] = m
    submission = importlib.import_module(root + ".flowformer.configs.submission")
    transformer = importlib.import_module(root + ".flowformer.core.FlowFormer.LatentCostFormer.transformer")

    def initialize_Flowformer():   # (9) no checkpoint load
        cfg = submission.get_cfg()
        return transformer.FlowFormer(cfg["latentcostformer"])

    pk.initialize_Flowformer = initialize_Flowformer
    gimmvfi_f = importlib.import_module(root + ".gimmvfi_f")
    _loaded.update(gimmvfi_f=gimmvfi_f, transformer=transformer, submission=submission)
    return _loaded


def build_reference_model_f(state_dict=None, seed=0):
    """The reference GIMMVFI_F (gimmvfi_f.py:27) on CPU, eval mode, random init (seeded) unless a state_dict is given."""
    mods = load_reference_f()
    torch.manual_seed(seed)
    cfg = default_arch_config()
    cfg["type"] = "gimmvfi_f"
    model = mods["gimmvfi_f"].GIMMVFI_F(cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model


if __name__ == "__main__":
    import time

    torch.set_grad_enabled(False)
    m = build_reference_model_f()
    sd = m.state_dict()
    print("GIMMVFI_F: %d tensors, %.2f M parameters" % (len(sd), sum(v.numel() for v in sd.values() if v.dtype.is_floating_point) / 1e6))
    H, W = 128, 160
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gimmvfi_b200.synth import synth_batch

    xs = synth_batch(1, H, W, seed=3)
    coord = [(m.sample_coord_input(1, (H, W), [0.5], device=xs.device), None)]
    t0 = time.time()
    out = m(xs, coord, t=[0.5 * torch.ones(1)])
    print("forward %dx%d: %.1f s; imgt_pred" % (H, W, time.time() - t0), tuple(out["imgt_pred"][0].shape), float(out["imgt_pred"][0].mean()))
