for v in default noinl noinl384 inl384; do
  if [ $v = default ]; then unset PCREG_LIB; else export PCREG_LIB=/root/repo/ab/libpcreg_$v.so; fi
  echo "== $v"; python tools/c4_check.py 16384 2>&1 | grep "profiling 2"
done
