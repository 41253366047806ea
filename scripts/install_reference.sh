#!/bin/bash
# One-off, offline "install" of the UNMODIFIED reference into baseline/_ref (git-ignored, travels with gpurun) so that
# `bench.py --impl reference` can time the reference's own functions on the GPU box's host cores.
# The reference has no setup.py / pyproject.toml (a flat directory of scripts), so
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails ("neither 'setup.py' nor 'pyproject.toml' found").  As the build notes allow, the install is done from a
# copy under /tmp to which ONLY a minimal setup.py is added (the reference's own files are untouched); the hot-path
# modules land in baseline/_ref/code/.  Nothing under baseline/_ref is tracked by git.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${1:-/root/reference}
[ -d "$SRC/code" ] || { echo "reference checkout not found at $SRC"; exit 1; }
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$SRC" \
  > /tmp/ref_install_plain.log 2>&1 && { echo "plain pip install worked"; exit 0; }
echo "plain pip install failed (expected: no setup.py): $(tail -1 /tmp/ref_install_plain.log)"
TMP=$(mktemp -d /tmp/yolo_ref_XXXX)
cp -r "$SRC/code" "$TMP/code"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup
setup(name="yolo-for-turbines-reference", version="0.0.0", packages=[],
      data_files=[("code", ["code/model.py", "code/utils.py", "code/loss.py", "code/config.py", "code/dataset.py"])])
PY
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP" \
  > /tmp/ref_install.log 2>&1 || { tail -5 /tmp/ref_install.log; exit 1; }
ls "$ROOT/baseline/_ref/code"
for f in model utils loss config dataset; do cmp "$SRC/code/$f.py" "$ROOT/baseline/_ref/code/$f.py"; done && echo "installed files are byte-identical to the reference"
