#!/bin/bash
# development aid: ms per erode call by batch size and launch shape (tools/variants.py builds; SHX_LIB selects one)
cd "$(dirname "$0")/.."
for v in "$@"; do
  export SHX_LIB=simplehydrology_b200/_variants/libshx_var_$v.so
  echo "== variant $v"
  python tools/tune_descend.py 16 0:0:0 2>&1 | grep mapsize
  python tools/tune_descend.py 1 0:0:0 2>&1 | grep mapsize
  python tools/tune_descend.py 4 0:0:0 448:3:0:1 64:2:0:1 2>&1 | grep mapsize
  for cy in 256 384 512; do python tools/tune_descend.py 8 0:0:0:0:$cy 448:3:0:1:$cy 64:2:0:1:$cy 448:6:0:0:$cy 2>&1 | grep mapsize; done
done
