#!/usr/bin/env bash
# Put the UNMODIFIED reference where the GPU box can see it: baseline/_ref/ is git-ignored
# (no reference source enters the history) but NOT gpurun-ignored, so it travels with the
# snapshot.  /root/reference exists only in the build container; on the GPU box nothing reads it.
#   tools/install_ref.sh [source-root]          (default: $BYZ_REFERENCE or /root/reference)
# The reference is plain Python scripts without setup.py / pyproject.toml, so `pip install
# --target baseline/_ref /root/reference` has nothing to build; the install is a verbatim copy
# (DESIGN.md §2 records this).  A manifest of sha256 sums is written next to it so that tests can
# assert the copy is byte-identical to the source.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="${1:-${BYZ_REFERENCE:-/root/reference}}"
dst="$here/baseline/_ref"
if [ ! -f "$src/aggregators/__init__.py" ]; then
  echo "install_ref: no reference at $src" >&2
  exit 1
fi
rm -rf "$dst"
mkdir -p "$dst"
# everything but VCS metadata, caches and the submodule's images
( cd "$src" && find . -type f -not -path './.git/*' -not -path '*/__pycache__/*' -not -name '*.png' -print0 \
  | xargs -0 -I{} cp --parents {} "$dst/" )
( cd "$dst" && find . -type f -not -name MANIFEST.sha256 -print0 | sort -z | xargs -0 sha256sum > MANIFEST.sha256 )
echo "install_ref: $(wc -l < "$dst/MANIFEST.sha256") files -> $dst"
