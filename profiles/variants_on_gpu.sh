#!/bin/bash
# One GPU call that measures every tuning build of the triangulation kernels and runs the parity suite on the candidates.
# Build the variants in the container first (nvcc cross-compiles; the .so files travel with the snapshot):
#     python profiles/build_variants.py
#     gpurun --timeout 600 -- 'bash profiles/variants_on_gpu.sh > gpurun_out/variants.log 2>&1'
# Nothing printed here is a bench value of record: a variant that wins becomes the default and bench.py measures it.
set -u
cd "$(dirname "$0")/.."
echo "== float storage, V = 8, clean input"
python profiles/tri_variant_check.py --dtype f32 --joints 17000000 --views 8
echo "== float storage, V = 8, 1 % unusable views"
python profiles/tri_variant_check.py --dtype f32 --joints 17000000 --views 8 --unusable 0.01
echo "== float storage, V = 16"
python profiles/tri_variant_check.py --dtype f32 --joints 10000000 --views 16
echo "== double storage, V = 8"
python profiles/tri_variant_check.py --dtype f64 --joints 17000000 --views 8
for v in ${CANDIDATES:-all lean_packed lean64}; do
    lib=profiles/variants/libmc3d_$v.so
    [ -f "$lib" ] || continue
    echo "== parity suite (triangulation + integration) with $v"
    MC3D_LIB=$PWD/$lib python -m pytest tests/test_triangulate_gpu.py tests/test_integration_stub_gpu.py -m gpu -q -x 2>&1 | tail -3
done
