set +e
for nt in 1000 5000; do
for v in t2 off t0 t1; do
  unset FWI_VARIANT_LIB; unset FWI_PDL
  [ $v = off ] && export FWI_PDL=0
  [ $v = t0 ] && export FWI_VARIANT_LIB=$PWD/build_variants/t0/libfwi_b200.so
  [ $v = t1 ] && export FWI_VARIANT_LIB=$PWD/build_variants/t1/libfwi_b200.so
  echo "== nt=$nt $v"; timeout 300 python tools/step_times.py $nt 2>&1 | tail -4
done
done
