#!/bin/bash
# No GPU needed: the `-m gpu` tests against the whole product library compiled for the host (tests/host_emul/fullhost,
# DESIGN 4.9).  Default: the device assertions that have not run on a B200 yet; `all` = the whole suite except the
# performance assertion and the 50 400-tet solve (minutes per kernel launch chain under emulation).
cd "$(dirname "$0")/.."
export PHB200_TEST_HOST_EMUL=1
if [ "$1" = "all" ]; then
  python -m pytest tests -q -m gpu -p no:cacheprovider --ignore tests/test_gpu_nccl.py \
    --deselect tests/test_gpu_parity.py::test_sumgat_and_fp64_peak \
    --deselect tests/test_phio.py::test_gpu_solve_from_files_c1_cube
else
  python -m pytest tests/test_zz_gpu_late.py -q -m gpu -p no:cacheprovider
fi
