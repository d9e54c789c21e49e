OUT=gpurun_out; mkdir -p $OUT
for lib in $1; do
for wl in c4 c3; do
PYCONTACT_B200_LIB=pycontact_b200/$lib timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu 2>$OUT/c34v.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$lib $wl value %.1f ms/step %.3f scanseq %.3f'%(j['value'],j['ms_per_step'],j['roofline'].get('scan_sequence_ms_per_batch',0)))
"
done
done
