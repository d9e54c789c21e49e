OUT=gpurun_out; mkdir -p $OUT
for s in 1 2 3 4; do
  timeout 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --slots $s 2>$OUT/c3x_slots$s.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('slots $s value %.0f ms/step %.3f scanseq %.3f kernel %.3f share %.3f'%(j['value'],j['ms_per_step'],j['roofline'].get('scan_sequence_ms_per_batch',0),j['roofline']['kernel_ms_per_launch'],j['roofline']['kernel_share_of_step']))
"
done
timeout 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --sparse-fold 2>$OUT/c3x_sparse.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('sparse value %.0f ms/step %.3f'%(j['value'],j['ms_per_step']), j['config'].get('results'))
"
timeout 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --batch 89 --slots 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('batch89 slots4 value %.0f ms/step %.3f'%(j['value'],j['ms_per_step']))
"
