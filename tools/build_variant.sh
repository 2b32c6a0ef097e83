#!/bin/bash
# tools/build_variant.sh NAME 'sed-expression' [file]   -> ab/libpcreg_NAME.so built from a patched copy of one source file
set -e
name=$1; expr=$2; file=${3:-nn_grid.cu}
cd /root/repo/pcreg_b200/csrc
cp $file /tmp/an/_orig_$file
sed -i "$expr" $file
if cmp -s $file /tmp/an/_orig_$file; then echo "variant $name: sed changed nothing"; exit 1; fi
cd /root/repo
python -c "
import sys; sys.path.insert(0,'.')
from pcreg_b200.build import build_library
build_library()" 2>&1 | tail -3
cp pcreg_b200/libpcreg_b200.so ab/libpcreg_$name.so
cp /tmp/an/_orig_$file pcreg_b200/csrc/$file
echo built ab/libpcreg_$name.so
