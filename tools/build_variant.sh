#!/bin/bash
# Builds a what-if variant of the library into build/variants/libfdf_<name>.so:  tools/build_variant.sh <name> <nvcc flags...>
# Run it on the GPU box with  FDF_LIB=build/variants/libfdf_<name>.so python tools/time_kernel.py
set -e
name=$1; shift
mkdir -p build/variants
make -s -C feature_detector_fast_b200/csrc -B OUT=../../build/variants LIBNAME=libfdf_$name.so EXTRA="$*" 2>&1 | grep -E "error|fdf_detect_kernelILi1ELi64|Used" | grep -A1 "ILi1ELi64" | tail -1
