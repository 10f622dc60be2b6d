#!/bin/bash
# build a kernel variant here (CPU box) into variants/<tag>.so : mkvariant.sh <tag> "<nvcc extra flags>"
tag=$1; shift
mkdir -p variants
bash scripts/variant.sh "$1" || exit 1
cp h264-scroll-encoder_b200/libh264scroll_b200.so variants/$tag.so
