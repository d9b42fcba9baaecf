set +e
for rep in 1 2; do
echo "== prefetch on"; unset FWI_VARIANT_LIB; timeout 200 python tools/jitter_check.py 2>&1 | sed -n 3,3p
echo "== prefetch off"; FWI_VARIANT_LIB=$PWD/build_variants/nopf/libfwi_b200.so timeout 200 python tools/jitter_check.py 2>&1 | sed -n 3,3p
done
timeout 900 python -m pytest tests/test_fd2d_gpu.py -q -x 2>&1 | tail -2
