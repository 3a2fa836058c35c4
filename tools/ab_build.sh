#!/bin/bash
# Measurement aid: builds a variant of the library with extra nvcc flags into tools/ab/<name>.so (git-ignored, travels with gpurun).
# usage: tools/ab_build.sh <name> [extra nvcc flags...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
make -s -j 16 -C "$ROOT/sdr-modem_b200" BUILD=/tmp/sdrm_ab_$name LIB="$ROOT/tools/ab/$name.so" EXTRA_NVCC="$*" 2>&1 | grep -v "deprecated-gpu-targets" || true
ls -la "$ROOT/tools/ab/$name.so"
