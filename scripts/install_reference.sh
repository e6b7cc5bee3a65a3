#!/bin/bash
# Offline install of the UNMODIFIED reference package into baseline/_ref (git-ignored, travels with gpurun).
# /root/reference declares the hatchling build backend, which is neither installed nor in /opt/wheelhouse, so the
# install runs from a copy under /tmp whose pyproject.toml names setuptools instead; the package directory itself
# is byte-identical (checked by the diff at the end).  --no-deps: einx is not available offline (oracle/einx_shim
# satisfies the import; the hot path never calls it).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/refcopy.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'EOF'
import sys
p = sys.argv[1]
s = open(p).read()
i = s.find('[tool.hatch')
if i >= 0:
    s = s[:i]
s = s.replace('requires = ["hatchling"]', 'requires = ["setuptools"]')
s = s.replace('build-backend = "hatchling.build"', 'build-backend = "setuptools.build_meta"')
s += '[tool.setuptools]\npackages = ["vector_quantize_pytorch"]\n'
open(p, 'w').write(s)
EOF
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP"
diff -rq -x __pycache__ "$SRC/vector_quantize_pytorch" "$ROOT/baseline/_ref/vector_quantize_pytorch"
rm -rf "$TMP"
echo "reference installed into $ROOT/baseline/_ref"
