#!/bin/bash
cd "$(dirname "$0")/.."
export SHX_LIB=simplehydrology_b200/_variants/libshx_var_n28.so
python tools/tune_descend.py 1 0:0:0 64:2:0:1 32:2:0:1 128:2:0:1 2>&1 | grep mapsize
python tools/tune_descend.py 2 0:0:0 64:2:0:1 2>&1 | grep mapsize
python tools/tune_descend.py 4 0:0:0 64:2:0:1 128:2:0:1 32:2:0:1 2>&1 | grep mapsize
python tools/tune_descend.py 8 64:2:0:1 128:2:0:1 32:2:0:1 2>&1 | grep mapsize
for cy in 256 512; do python tools/tune_descend.py 16 448:3:0:1:$cy 64:2:0:1:$cy 128:2:0:1:$cy 96:2:0:1:$cy 2>&1 | grep mapsize; done
