#!/bin/bash
# Installs the UNMODIFIED reference (LibKGE fork, /root/reference) into the git-ignored baseline/_ref/ so that it
# travels to the GPU box with the gpurun snapshot (bench.py --impl reference, tests/test_gpu_dropin.py).
# Build container only: /root/reference does not exist on the GPU box.
#   1. the prescribed offline pip install (from a copy: the build writes into the source tree);
#   2. the reference's setup.py lists packages=["kge"] only -- it is meant to be installed in develop mode
#      (`pip install -e .`) -- so the wheel lacks the sub-packages (kge/job, kge/model, kge/util) and every yaml file;
#      they are copied verbatim on top.  Nothing is edited.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${1:-/root/reference}
[ -d "$REF/kge" ] || { echo "no reference tree at $REF"; exit 0; }
rm -rf /tmp/ref_copy "$ROOT/baseline/_ref"
cp -r "$REF" /tmp/ref_copy
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$ROOT/baseline/_ref" /tmp/ref_copy 2>&1 | tail -2
cp -r "$REF/kge/." "$ROOT/baseline/_ref/kge/"
mkdir -p "$ROOT/baseline/_ref/examples" && cp "$REF"/examples/*.yaml "$ROOT/baseline/_ref/examples/"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
diff -rq "$REF/kge" "$ROOT/baseline/_ref/kge" && echo "baseline/_ref/kge identical to $REF/kge"
