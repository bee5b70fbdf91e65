#!/usr/bin/env bash
# Installs the UNMODIFIED reference (csxeba/brainforge, pure Python) from /root/reference into oracle/_ref/ so that
# bench.py --impl reference (and the cpu_baseline leg) can time the reference's own NumPy / Numba code path on the GPU
# box's host cores.  oracle/_ref/ is git-ignored (no reference source enters the history) but travels with the repo
# snapshot to the GPU box.  Authoring container only: /root/reference does not exist on the GPU box.
#
#   bash oracle/build_ref.sh
#
# pip needs a writable source tree (setup.py writes egg-info), hence the copy under /tmp.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/brainforge" ]; then
  echo "build_ref: $SRC not found -- keeping whatever oracle/_ref already holds" >&2
  exit 0
fi
TMP="$(mktemp -d /tmp/bf_ref_XXXX)"
cp -r "$SRC"/. "$TMP"/
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" \
  || { echo "build_ref: pip install failed, copying the package directory instead" >&2; mkdir -p "$HERE/_ref"; cp -r "$SRC/brainforge" "$HERE/_ref/brainforge"; }
rm -rf "$TMP"
python - <<PY
import os, sys, tempfile
os.environ["HOME"] = tempfile.mkdtemp(prefix="bfhome_")
sys.path.insert(0, os.path.join("$HERE", "_ref"))
import numpy as np
np.asscalar = lambda a: a.item()
import brainforge
print("oracle/_ref: brainforge importable from", os.path.dirname(brainforge.__file__))
PY
