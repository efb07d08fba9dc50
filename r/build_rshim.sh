#!/bin/sh
# Build the R .Call shim on a machine that has R (this image has none: the shim has never been compiled here).
#   sh r/build_rshim.sh        -> r/potus_b200_rshim.so ; then in R:  source("r/potus_b200.R")
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
python3 -c "import sys; sys.path.insert(0, '$ROOT'); import __graft_entry__ as g; g.build()"   # nvcc: us-potus-model_b200/lib/libpotus_b200.so
cd "$ROOT/r"
PKG_CPPFLAGS="-I$ROOT/include" PKG_LIBS="-L$ROOT/us-potus-model_b200/lib -lpotus_b200 -Wl,-rpath,$ROOT/us-potus-model_b200/lib" \
  R CMD SHLIB -o potus_b200_rshim.so potus_b200_rshim.c
echo "built $ROOT/r/potus_b200_rshim.so"
