#!/bin/bash
# Place the reference sources the drop-in test needs (tests/test_gpu_parity.py::test_unmodified_reference_drivers_on_native_shells)
# under tests/_reference_payload/ so that a gpurun call carries them to the GPU box as TEST DATA.  The directory is
# git-ignored: reference sources never enter this repository's history.  Run in the build container only.
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_reference_payload
cp /root/reference/orca_predict.py /root/reference/orca_modules.py /root/reference/orca_models.py tests/_reference_payload/
echo "shipped: $(ls tests/_reference_payload | tr '\n' ' ')"
