"""Installs the UNMODIFIED reference package into baseline/_ref/ (git-ignored; it travels to the GPU box).

    python baseline/install_reference.py

`pip install --no-index --find-links /opt/wheelhouse --target baseline/_ref /root/reference` fails in this image: the
reference's build backend (hatchling, pyproject.toml:1-3) is not in the wheelhouse.  Its wheel target is
`packages = ["most_queue"]` (pyproject.toml:74-75) and the package is pure Python, so what pip would have written
into the target directory is exactly a copy of the `most_queue/` tree — which is what this script does, byte for byte,
followed by a hash check of every file against its source.  The only third-party import the hot path needs and the
image lacks is `colorama` (ANSI colour strings, sim/base.py:7); a six-line stand-in is written next to the package
(NOT into it).  Nothing under baseline/_ref/ is edited.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MQ_REFERENCE", "/root/reference")

COLORAMA_STUB = '''"""Stand-in for the `colorama` package (absent from this image): every colour code is the empty string."""


class _Blank:
    def __getattr__(self, _):
        return ""


Fore = Style = Back = _Blank()


def init(*a, **k):
    return None
'''


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(verbose: bool = False) -> str | None:
    """Returns the install directory, or None when the reference tree is not present (the GPU box)."""
    src_pkg = os.path.join(SRC, "most_queue")
    if not os.path.isdir(src_pkg):
        return DST if os.path.isdir(os.path.join(DST, "most_queue")) else None
    dst_pkg = os.path.join(DST, "most_queue")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = 0
    for root, _, files in os.walk(src_pkg):
        if "__pycache__" in root:
            continue
        for f in files:
            if f.endswith(".pyc"):
                continue
            a = os.path.join(root, f)
            b = os.path.join(dst_pkg, os.path.relpath(a, src_pkg))
            if _sha(a) != _sha(b):
                raise RuntimeError(f"copy of {a} differs")
            n += 1
    try:
        import colorama  # noqa: F401
    except ImportError:
        with open(os.path.join(DST, "colorama.py"), "w") as f:
            f.write(COLORAMA_STUB)
    with open(os.path.join(DST, "INSTALLED_FROM.txt"), "w") as f:
        f.write(f"{SRC}/most_queue ({n} files, sha256-verified copy; pyproject version 2.9)\n")
    if verbose:
        print(f"reference installed into {DST}: {n} files")
    return DST


def add_to_path() -> bool:
    """Makes `import most_queue` resolve to baseline/_ref; False when it is not installed."""
    if not os.path.isdir(os.path.join(DST, "most_queue")):
        return False
    if DST not in sys.path:
        sys.path.insert(0, DST)
    return True


if __name__ == "__main__":
    print(install(verbose=True))
