#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Compiles the UNMODIFIED reference sources where they
# lie (default /root/reference/src/*.cpp; never copied into this repo) together with
# oracle/ref_driver.cpp into oracle/_ref/libcafe_ref.so.  oracle/_ref/ is git-ignored but travels to
# the GPU box with the snapshot.  Does not use the reference's own CMake build; config.h is
# generated here with the constants CMakeLists.txt:10-15,33 would configure.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CAFE_REF_DIR:-/root/reference}"
OUT="$HERE/_ref"
CXX="${CAFE_REF_CXX:-/usr/bin/g++}"   # not $CXX: the image sets it to a wrapper without libgomp.spec
if [ ! -d "$REF/src" ]; then
  echo "reference not present at $REF; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
cat > "$OUT/config.h" <<'CFG'
#define PROJECT_NAME "CAFE"
#define PROJECT_VER  "1.1"
#define PROJECT_VER_MAJOR "1"
#define PROJECT_VER_MINOR "1"
#define PROJECT_VER_PATCH ""
#define BOUNDING_STEP_SIZE 20
#define MATRIX_SIZE_MULTIPLIER 3.0
#define MATRIX_EPSILON 5e-6
#define OPTIMIZER_HIGH_PRECISION 1e-6
#define OPTIMIZER_LOW_PRECISION  1e-3
#define PHASED_OPTIMIZER_PHASE1_ATTEMPTS  4
#define NUM_OPTIMIZER_INITIALIZATION_ATTEMPTS  1
#define LAMBDA_PERTURBATION_STEP_SIZE  1
#define HAVE_GETOPT_H 1
#define LOG_OFFSET 1.0
#define TRANSCRIPT_RECONSTRUCTION 2
#define SIMULATOR 3
#define MATRIX 6
#define INFERENCE 8
CFG
FLAGS=(-std=c++11 -O2 -fopenmp -fPIC -w -fno-access-control -include "$OUT/config.h" -I"$REF/src" -I"$HERE/../include"
       -DDOCTEST_CONFIG_DISABLE -DSILENT -DELPP_NO_CHECK_MACROS -DMODEL_GENE_EXPRESSION_LOGS
       -DOPTIMIZER_STRATEGY=NelderMead -DDISCRETIZATION_RANGE=200 -DMAX_STACK_FAMILY_SIZE=1000)
objs=()
pids=()
SHIM_OBJ="$OUT/obj/ref_gpu_model.o"
for src in "$REF"/src/*.cpp "$HERE/ref_driver.cpp" "$HERE/ref_gpu_model.cpp"; do
  [ -f "$src" ] || continue
  obj="$OUT/obj/$(basename "${src%.cpp}").o"
  [ "$obj" = "$SHIM_OBJ" ] || objs+=("$obj")
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/../cafe5_b200/host/gpu_model.hpp" -nt "$obj" ] || [ "$HERE/../include/cafe_b200.h" -nt "$obj" ] \
     || [ "$HERE/ref_optimize.hpp" -nt "$obj" ] || [ "$HERE/ref_ctx.hpp" -nt "$obj" ]; then
    "$CXX" "${FLAGS[@]}" -I"$HERE" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do [ -n "$p" ] && { wait "$p" || fail=1; }; done
[ "$fail" = 0 ] || { echo "compilation failed" >&2; exit 1; }
# 1. the unmodified reference + its driver: no product code inside, nothing of the product linked
"$CXX" -shared -fopenmp -o "$OUT/libcafe_ref.so" "${objs[@]}" -lz -ldl
echo "built $OUT/libcafe_ref.so"
# 2. the drop-in shim's test driver binds the product's C ABI: its own library, on top of (1) and libcafe_b200.so
if [ -f "$HERE/../cafe5_b200/libcafe_b200.so" ]; then
  "$CXX" -shared -fopenmp -o "$OUT/libcafe_ref_shim.so" "$SHIM_OBJ" -L"$OUT" -lcafe_ref -L"$HERE/../cafe5_b200" -lcafe_b200 \
      '-Wl,-rpath,$ORIGIN' '-Wl,-rpath,$ORIGIN/../../cafe5_b200'
  echo "built $OUT/libcafe_ref_shim.so"
else
  echo "libcafe_b200.so not built yet: libcafe_ref_shim.so skipped (python -c 'import __graft_entry__ as g; g.build()')" >&2
fi
