#!/bin/bash
# Places the UNMODIFIED reference modules the full-pipeline checks import (BASELINE configs[2]: new NeRF branch + reference
# decoder) under oracle/_ref/ -- git-ignored, so the sources never enter the history, but not gpurun-ignored, so the copy
# travels to the GPU box where /root/reference does not exist.  Run by __graft_entry__.build() whenever the reference
# checkout is present.  Test / bench infrastructure only: the product never imports anything from here.
#   exp/cips3d/models/model_v3.py   Generator + Decoder (the caller of the NeRF branch, and the 2-D decoder)
#   exp/cips3d/volume_renderer.py   VolumeFeatureRenderer (the module the product replaces: the comparison arm)
#   exp/cips3d/nerf_utils.py        Render / Camera helpers the generator calls
set -e
SRC=${C3D_REFERENCE_SRC:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
DST=$HERE/_ref
[ -d "$SRC/exp/cips3d" ] || { echo "make_ref: no reference checkout at $SRC"; exit 0; }
mkdir -p "$DST/exp/cips3d/models" "$DST/exp/stylesdf"
cp "$SRC/exp/cips3d/models/model_v3.py" "$DST/exp/cips3d/models/"
cp "$SRC/exp/cips3d/volume_renderer.py" "$SRC/exp/cips3d/nerf_utils.py" "$DST/exp/cips3d/"
echo "make_ref: reference modules placed under $DST"
