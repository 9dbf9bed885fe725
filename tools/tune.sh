#!/bin/bash
# usage: tools/tune.sh <file.cu> "<defs variant 1>" "<defs variant 2>" ...   (run on the GPU box)
f=$1; shift
for v in "$@"; do
  touch nnnoiseless_b200/csrc/$f
  NNB_EXTRA_NVCC="$v" python -m nnnoiseless_b200.build > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py --streams 65536 --frames 8 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '%.3g'%d['value'], {k:round(x['ms'],3) for k,x in d['roofline']['kernels'].items()})"
done
touch nnnoiseless_b200/csrc/$f; python -m nnnoiseless_b200.build > /dev/null 2>&1
