#!/bin/bash
# rebuild the host library + all kernel images, print per-app registers and the packed-op counts
set -e
R=/root/repo
make -C $R/shaderbox_b200/csrc -j8 CXX=g++ 2>&1 | grep -i -A8 "error" | head -40 || true
make -C $R/shaderbox_b200/csrc plugins CXX=g++ 2>&1 | grep -v "^{\|sbx_cli compile" | tail -30 || true
for a in ${@:-CLOUDS PLANET}; do
  f=$R/shaderbox_b200/images/APP_$a.plugin.cubin
  cuobjdump -sass $f > /tmp/$a.sass
  echo "$a $(cuobjdump -res-usage $f | grep -o 'REG:[0-9]*') instrs=$(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/$a.sass) packed=$(grep -c 'FFMA2\|FMUL2\|FADD2' /tmp/$a.sass) LDG=$(grep -c 'LDG' /tmp/$a.sass)"
done
