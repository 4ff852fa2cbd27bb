#!/bin/bash
# Run a command inside .frozen/<tag> (see freeze.sh) with its gpurun_out/ pointing at the real one.
TAG=$1; shift
cd "$(dirname "$0")/../../.frozen/$TAG" || exit 9
rm -rf gpurun_out; ln -s ../../gpurun_out gpurun_out
bash -c "$*"
