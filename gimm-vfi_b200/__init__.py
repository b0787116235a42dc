"""gimmvfi_b200 — B200-native (sm_100a) inference path for GIMM-VFI's per-pair
interpolation, drop-in for ``GIMMVFI_R.forward`` of GSeanCDAT/GIMM-VFI."""
__version__ = "0.1.0"

from .config import ConfigNode, default_arch_config, load_config  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require torch.cuda or the .so
    if name in ("GIMMVFI_R", "create_model", "sample_coords"):
        from . import model

        return getattr(model, name)
    if name == "GIMM":
        from .gimm import GIMM

        return GIMM
    if name == "GIMMVFI_F":
        from .model_f import GIMMVFI_F

        return GIMMVFI_F
    if name == "EngineHandle":
        from .engine import EngineHandle

        return EngineHandle
    raise AttributeError(name)
