#!/bin/bash
# Static SASS statistics of one tile-kernel instantiation (default: float, VEC 4, hybrid, z and t on, TSMODE 0) of a library build.
# usage: scripts/sass_tile.sh [lib.so] [mangled-substring]
LIB=${1:-pytv-4d_b200/csrc/libpytv_b200.so}
PAT=${2:-tv_tile_kernelIfLi4ELi3ELb1ELb1ELi4ELi0E}
cuobjdump -sass $LIB 2>/dev/null | awk -v pat="$PAT" '
  /Function :/ { on = index($0, pat) > 0 }
  on && /^ +\/\*[0-9a-f]+\*\/ / { n++; op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); c[op]++ }
  END { printf "total %d\n", n; for (k in c) printf "%6d %s\n", c[k], k | "sort -nr | head -22" }'
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 "$PAT" | grep REG
