#!/bin/bash
# Places the UNMODIFIED reference package (dobraczka/kiez v0.5.0, pure Python) under
# baseline/_ref/ so that `bench.py --impl reference` can drive the reference's own
# Kiez(SklearnNN brute, hubness=...) on the GPU box, where /root/reference does not exist.
# baseline/_ref/ is git-ignored (never committed) but travels with the gpurun snapshot.
# `pip install --target baseline/_ref /root/reference` is tried first (the base contract's
# recipe); it fails here because the build backend (poetry-core) is not in the offline
# wheelhouse, so the package directory is copied as is -- the same files pip would install.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${KIEZ_REFERENCE_ROOT:-/root/reference}"
DEST="$ROOT/baseline/_ref"
if [ ! -d "$REF/kiez" ]; then
  echo "vendor_reference: $REF/kiez not found (nothing to do on the GPU box)"; exit 0
fi
mkdir -p "$DEST"
if [ "${KB2_TRY_PIP:-0}" = "1" ] && python -m pip install --no-index --no-build-isolation --no-deps \
     --find-links /opt/wheelhouse --target "$DEST" "$REF" >/dev/null 2>"$DEST/pip_install.log"; then
  echo "vendor_reference: pip-installed the reference into $DEST"
else
  rm -rf "$DEST/kiez"
  cp -r "$REF/kiez" "$DEST/kiez"
  find "$DEST/kiez" -name '__pycache__' -type d -prune -exec rm -rf {} +
  echo "vendor_reference: copied $REF/kiez -> $DEST/kiez ($(find "$DEST/kiez" -name '*.py' | wc -l) files)"
fi
