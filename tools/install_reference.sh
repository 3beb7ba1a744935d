#!/bin/sh
# Copies the UNMODIFIED reference sources (read-only checkout /root/reference) into the git-ignored baseline/_ref/, which
# travels to the GPU box with the repo snapshot (gpurun) but never enters the history.  Needed only for
#   * the secondary baseline: the reference's own PyTorch path on the B200 (tools/bench_reference_gpu.py), and
#   * tests marked `reference` that run the reference's own step methods against this library on the GPU.
# The reference is pure Python (no build step, no setup.py); `pip install --target baseline/_ref /root/reference` has
# nothing to install, so this is a plain copy.
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
mkdir -p baseline/_ref
cp -r /root/reference/src baseline/_ref/src
cp -r /root/reference/configs baseline/_ref/configs
echo "installed $(find baseline/_ref -name '*.py' | wc -l) reference modules under baseline/_ref"
