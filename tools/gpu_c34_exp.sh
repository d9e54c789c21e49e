OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "synthetic or fused or edge or sparse or packed or cutoff or golden" > $OUT/c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/c34_pytest.log
for wl in c4 c3; do
timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu 2>$OUT/c34.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$wl value %.1f ms/step %.3f scanseq %.3f evals/s %.3g contacts %d'%(j['value'],j['ms_per_step'],j['roofline'].get('scan_sequence_ms_per_batch',0),j['roofline']['pair_evals_per_s'],j['per_step']['contacts']))
"
done
