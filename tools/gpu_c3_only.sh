OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "synthetic or fused or packed or sparse" > $OUT/c3o_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/c3o_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu 2>$OUT/c3o.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('c3 value %.1f ms/step %.3f scanseq %.3f'%(j['value'],j['ms_per_step'],j['roofline'].get('scan_sequence_ms_per_batch',0)))
"
done
