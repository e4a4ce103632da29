mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "tail_split or batch_rows or bottleneck_layer" > gpurun_out/r2_split_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_split_t.log
tail -3 gpurun_out/r2_split_t.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity"
for v in 0 1; do
  RNB_NO_SPLIT=$v $B > gpurun_out/r2_split_r50_ns$v.json 2> gpurun_out/r2_split_r50_ns$v.err
  RNB_NO_SPLIT=$v $B --arch resnet152 --batch 128 > gpurun_out/r2_split_r152_ns$v.json 2> gpurun_out/r2_split_r152_ns$v.err
  RNB_NO_SPLIT=$v RNB_C3N1=1 $B --arch resnet152 --batch 128 > gpurun_out/r2_split_r152_c3n1_1_ns$v.json 2> gpurun_out/r2_split_r152_c3n1_1_ns$v.err
done
for f in gpurun_out/r2_split_r*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('sustained',{}).get('ms_per_step'), d['clocks'])
"; done
