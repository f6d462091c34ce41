"""Place an UNMODIFIED copy of the reference under baseline/_ref/ (git-ignored; it travels to the GPU box with the snapshot).

The reference is plain Python without setup.py / pyproject.toml, so the `pip install --target baseline/_ref` of the bench contract
has nothing to build: the "install" is a file copy.  Used by
  * bench.py --impl reference      (the reference's own build_model(cfg) + model.eval()(NestedTensor) on the host cores),
  * tests/test_eval_loop_gpu.py    (the reference's deploy_model + validate_tuber_detection driving this repo's model).
__graft_entry__.build() calls this whenever /root/reference is present.  Nothing under baseline/_ref is imported by the product path.

    python tools/install_reference.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def install(src: str = "/root/reference", dest: str = DEST) -> str | None:
    if not os.path.isdir(src):
        return None
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    shutil.copytree(src, dest, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    for root, dirs, files in os.walk(dest):                      # the source mount is read-only; the copy need not be
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("TUBER_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    print(install(a.src) or f"{a.src} not found: nothing installed")
