#!/usr/bin/env bash
# Installs the UNMODIFIED reference (keurfonluu/stochopy, /root/reference) into the
# git-ignored baseline/_ref/ so that `bench.py --impl reference` and the cpu_baseline
# leg can import it on the GPU box (baseline/_ref travels with the gpurun snapshot;
# /root/reference does not exist there).  The reference is pure Python: nothing is
# compiled.  /root/reference is read-only and the build writes an egg-info into the
# source tree, so the install runs from a copy under /tmp; dependency resolution is
# skipped (--no-deps: importlib_metadata is not in the offline wheelhouse as a wheel
# but is importable in the image, as are numpy and joblib).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
src="${1:-/root/reference}"
if [ ! -d "$src/stochopy" ]; then
  echo "fetch_ref: $src/stochopy not found (nothing to install)" >&2
  exit 0
fi
tmp="$(mktemp -d /tmp/stochopy_ref.XXXXXX)"
cp -r "$src/." "$tmp/"
rm -rf "$here/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps \
  --find-links /opt/wheelhouse --target "$here/_ref" "$tmp"
rm -rf "$tmp"
python - <<PY
import sys
sys.path.insert(0, "$here/_ref")
import stochopy
print("fetch_ref: installed stochopy", stochopy.__version__, "->", stochopy.__file__)
PY
