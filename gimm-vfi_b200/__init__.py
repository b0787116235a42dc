"""gimmvfi_b200 — B200-native inference path for GIMM-VFI's per-pair interpolation."""
__version__ = "0.1.0"
