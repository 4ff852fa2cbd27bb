#!/bin/bash
# Freeze the working tree (sources + built .so) under .frozen/<tag>/ so that a queued gpurun call -- which snapshots
# /root/repo only when a GPU slot frees up, possibly minutes later -- runs exactly this state while editing continues.
# Usage: tools/gpu/freeze.sh <tag>; then run commands as:  tools/gpu/in_frozen.sh <tag> '<command>'
set -e
TAG=$1
cd "$(dirname "$0")/../.."
rm -rf .frozen
mkdir -p .frozen/$TAG
tar --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude=./.pytest_cache --exclude='__pycache__' \
    --exclude='./tools/cpu_emul/_emul_gen.cpp' --exclude='./tools/cpu_emul/*.so' --exclude='./tools/ubench' -cf - . | tar -xf - -C .frozen/$TAG
echo "frozen -> .frozen/$TAG ($(du -sh .frozen/$TAG | cut -f1))"
