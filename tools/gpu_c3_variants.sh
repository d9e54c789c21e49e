# usage: tools/gpu_c3_variants.sh "lib1 lib2 ..." "extra bench args"
OUT=gpurun_out; mkdir -p $OUT
for lib in $1; do
  for s in 3 4; do
  PYCONTACT_B200_LIB=pycontact_b200/$lib timeout 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --slots $s $2 2>$OUT/c3v.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$lib slots $s value %.0f ms/step %.3f scanseq %.3f kernel %.3f share %.3f'%(j['value'],j['ms_per_step'],j['roofline'].get('scan_sequence_ms_per_batch',0),j['roofline']['kernel_ms_per_launch'],j['roofline']['kernel_share_of_step']))
"
  done
done
