OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "synthetic or fused or edge or sparse or packed or self or golden or cutoff" > $OUT/c34t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/c34t_pytest.log
bash tools/gpu_c34_variants.sh "libpycontact_b200.so"
