#!/usr/bin/env bash
# Makes the UNMODIFIED reference forward importable on the GPU box: copies the four files its forward needs
# (model.py -> layers.py, utility_layers.py, training_utilities.py) from /root/reference into the git-ignored
# baseline/_ref/ (the reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
# has nothing to install; gpurun ships baseline/_ref/ like the built .so).  Nothing here is committed or edited.
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC" ]; then
  echo "vendor_reference: $SRC not present (GPU box?) -- keeping what is in $DST" >&2
  exit 0
fi
mkdir -p "$DST"
for f in model.py layers.py utility_layers.py training_utilities.py model_config_vit.yaml; do
  cp "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum model.py layers.py utility_layers.py training_utilities.py model_config_vit.yaml ) > "$DST/SHA256SUMS"
echo "vendored $(wc -l < "$DST/SHA256SUMS") reference files into $DST"
