#!/usr/bin/env python
"""Stage the UNMODIFIED reference package into the git-ignored baseline/_ref/ so that it travels to the GPU box.

    python scripts/stage_reference.py            # needs /root/reference (build container only)

`python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref
/tmp/copy-of-reference` was tried first and fails: the project's build backend is poetry-core, which is neither
installed nor in /opt/wheelhouse. The package is pure Python, so what pip would have put under --target is exactly
the `ocrs_models/` directory; this script copies it byte for byte (plus the licence files) and records a manifest of
sha256 sums so a reader can check that nothing was edited. baseline/_ref/ is listed in .gitignore: no reference
source enters the history."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("OCRS_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main() -> int:
    pkg = os.path.join(SRC, "ocrs_models")
    if not os.path.isdir(pkg):
        print(f"stage_reference: {pkg} not found (GPU box?) - keeping whatever is in {DST}")
        return 0 if os.path.isdir(os.path.join(DST, "ocrs_models")) else 1
    shutil.rmtree(os.path.join(DST, "ocrs_models"), ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(pkg, os.path.join(DST, "ocrs_models"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for d, _, files in os.walk(os.path.join(DST, "ocrs_models")):
        for f in sorted(files):
            p = os.path.join(d, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"staged {len(manifest)} files from {pkg} into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
