#!/usr/bin/env bash
# Builds the TensorFlow op library over libm4d.  Needs a Python with TensorFlow 2.x (the reference pins tensorflow-gpu 2.7,
# README.md:79); this repository's image has none, so nothing here runs in its CI.
#   ./build.sh [output dir]     default: <reference>/utils/special_ops - where utils/dense_image_warp.py:38 looks
# (the reference's own make.sh:10 writes cuda_backproject/backproject.so, which that loader never finds.)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="${1:-$PWD/utils/special_ops}"
python - <<'PY' || { echo "TensorFlow is not importable: nothing to build" >&2; exit 1; }
import tensorflow  # noqa: F401
PY
TF_CFLAGS=( $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))') )
TF_LFLAGS=( $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))') )
python "$ROOT/m4depth_b200/_build.py"                      # libm4d.so (nvcc, sm_100a)
mkdir -p "$OUT"
g++ -std=c++17 -shared -fPIC -O2 "$HERE/m4d_tf_ops.cc" -o "$OUT/backproject.so" \
    "${TF_CFLAGS[@]}" -DGOOGLE_CUDA=1 -I/usr/local/cuda/include \
    -L"$ROOT/m4depth_b200" -l:libm4d.so -Wl,-rpath,"$ROOT/m4depth_b200" "${TF_LFLAGS[@]}"
echo "wrote $OUT/backproject.so (ops: BackProject, BackProjectGrad, M4dPscvFused, M4dPscvFusedGrad, M4dSncv)"
