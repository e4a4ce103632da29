# 8-GPU confirmation of the final build: the C++ multi-GPU driver, the torchrun bench (configs[3]: 2048 images as 8 x 256),
# ResNet-152 1024 images over 8 GPUs (configs[4])
mkdir -p gpurun_out
W=$(python -c "from resnet_c_b200 import weights; print(weights.cached_weights_dir('resnet50', 0))")
./build/resnet_infer_mgpu resnet50 bf16 8 256 20 $W > gpurun_out/mgpu8b.txt 2> gpurun_out/mgpu8b.err; echo "mgpu8 rc=$?"; head -1 gpurun_out/mgpu8b.txt; tail -2 gpurun_out/mgpu8b.err
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 20 --warmup 5 --scaling strong --global-batch 2048 --no-cpu-baseline > gpurun_out/bench8b_strong2048.json 2> gpurun_out/bench8b_strong2048.err; echo "bench8 rc=$?"
$T bench.py --gpus 8 --steps 20 --warmup 5 --arch resnet152 --batch 128 --no-cpu-baseline > gpurun_out/bench8b_r152.json 2> gpurun_out/bench8b_r152.err; echo "bench8 r152 rc=$?"
for f in gpurun_out/bench8b_strong2048.json gpurun_out/bench8b_r152.json; do python - <<PY
import json
d=json.loads(open("$f").read().strip().splitlines()[-1])
print("$f", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "u8", round(d["e2e_u8"]["value"]), "gather_ok", d["gather_ok"], "parity", d["parity"]["ok"], d["parity"]["rel_err"], "frac", d["roofline"]["frac"], d["sustained"].get("ms_per_step"))
PY
done
