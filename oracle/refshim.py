"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Import shim for running the *unmodified* reference modules from
/root/reference in this container (they are read, never copied): ``download_data.py:14-15`` imports netCDF4,
which is not installed and is not needed by the hot path, so an empty stand-in module is registered first
(SURVEY §8-c)."""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WINDSR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CNN_models"))


def activate():
    """Make ``CNN_models`` / ``GAN_models`` / ``process_data`` / ``config.config`` of the reference importable."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if "netCDF4" not in sys.modules:
        m = types.ModuleType("netCDF4")
        m.Dataset = m.MFDataset = object
        sys.modules["netCDF4"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
