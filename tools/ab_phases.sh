#!/bin/bash
# A/B of the per-phase times of the persistent loop: ADMMQ_LIB=<other build> against the in-tree library.
# usage: tools/ab_phases.sh [libA libB ...]   (default: lib/libadmmq_prev.so lib/libadmmq.so)
cd "$(dirname "$0")/.."
LIBS=${@:-"admm-quantization_b200/lib/libadmmq_prev.so admm-quantization_b200/lib/libadmmq.so"}
for lib in $LIBS; do
  [ -f "$lib" ] || continue
  echo "=== $lib"
  for spec in "layer4.1.conv1 32 0" "layer4.1.conv1 36 0" "layer4.1.conv1 148 0" "layer4.1.conv1 36 2" "layer4.1.conv1 32 2" \
              "layer3.1.conv1 7 0" "layer3.1.conv1 7 2" "layer2.1.conv1 2 0" "layer2.1.conv1 2 2" "layer1.0.conv1 1 0" "layer1.0.conv1 1 2"; do
    set -- $spec
    ADMMQ_LIB=$PWD/$lib python tools/profile_target.py 300 $1 $2 1 $3
  done
done
