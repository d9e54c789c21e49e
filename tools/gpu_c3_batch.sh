OUT=gpurun_out; mkdir -p $OUT
for cfg in "--frames 712 --batch 178" "--frames 740 --batch 148" "--frames 740 --batch 148 --slots 4" "--frames 592 --batch 296 --slots 2"; do
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu $cfg 2>$OUT/c3b.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$cfg value %.1f ms/step %.3f scanseq %.3f step_gbs %.0f'%(j['value'],j['ms_per_step'],j['roofline'].get('scan_sequence_ms_per_batch',0),j['roofline'].get('step_gbs',0)))
"
done
