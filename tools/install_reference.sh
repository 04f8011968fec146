#!/bin/bash
# Installs the UNMODIFIED reference (pnnl/nmrfit at /root/reference, read-only) into the git-ignored baseline/_ref so
# that it travels to the GPU box with gpurun.  Only tests/test_gpu_reference_dropin.py uses it (skipped when absent);
# nothing in the product, bench.py or smoke() reads it.  The build wants to write into the source tree, hence the copy.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/nmrfit_refcopy baseline/_ref
cp -r /root/reference /tmp/nmrfit_refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/nmrfit_refcopy
ls baseline/_ref
