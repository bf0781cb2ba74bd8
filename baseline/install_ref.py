"""Installs the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that ``bench.py --impl reference``
can drive the stock ``wind_field_GAN_3D.optimize_parameters`` on the GPU box's host cores (``/root/reference`` does not
exist there; ``baseline/_ref/`` travels with the gpurun snapshot, it is NOT in .gpurunignore).

``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference``
was tried first and cannot work: the project's build backend is poetry-core (not in the image, no network) and its
pyproject names a package directory ``gan_sr_wind_field`` that does not exist in the repository — the reference is a
flat collection of scripts.  So this script copies exactly the files the training step imports:
``CNN_models/ GAN_models/ tools/ config/ process_data.py download_data.py`` and the shipped ``pretrained_models/*/
config.ini``.  Nothing under ``baseline/_ref`` is ever committed (.gitignore) and nothing in the product imports it.

    python baseline/install_ref.py [--src /root/reference]
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
ITEMS = ["CNN_models", "GAN_models", "tools", "config", "process_data.py", "download_data.py", "pretrained_models",
         "LICENSE.txt"]


def install(src="/root/reference", quiet=False) -> bool:
    if not os.path.isdir(os.path.join(src, "CNN_models")):
        if not quiet:
            print(f"install_ref: {src} not found (nothing installed)", file=sys.stderr)
        return False
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        s, d = os.path.join(src, item), os.path.join(DST, item)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pth"))
        elif os.path.exists(s):
            shutil.copy2(s, d)
    if not quiet:
        print(f"install_ref: reference files copied to {DST}")
    return True


def available() -> bool:
    return os.path.isdir(os.path.join(DST, "CNN_models"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    install(ap.parse_args().src)
