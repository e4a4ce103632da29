#!/bin/bash
# (8x the box time is charged: keep both runs short) 8 GPUs, one process per GPU: the torchrun bench with the host-packing decision made under contention (every rank at
# once), and the same with RNB_HOST_PACK=0 (plain FP32 copies) for comparison
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
RNB_VERBOSE=1 timeout 400 $T bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8_hp_auto.json 2> gpurun_out/bench8_hp_auto.err; echo "auto rc=$?"
RNB_HOST_PACK=0 timeout 400 $T bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8_hp_off.json 2> gpurun_out/bench8_hp_off.err; echo "off rc=$?"
grep "host pack" gpurun_out/bench8_hp_auto.err | head -8
for f in gpurun_out/bench8_hp_auto.json gpurun_out/bench8_hp_off.json; do python - <<PY
import json
d=json.loads(open("$f").read().strip().splitlines()[-1])
print("$f", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("fp32_copy_value"), d["e2e"].get("host_pack"), "u8", round(d["e2e_u8"]["value"]), "gather_ok", d["gather_ok"], "parity", d["parity"]["ok"])
PY
done
