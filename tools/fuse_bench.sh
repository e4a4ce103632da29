#!/bin/bash
# usage: tools/fuse_bench.sh tag "fuse next" ...   (GPU box) — bench + per-launch profile per config
tag=$1; shift
for cfg in "$@"; do set -- $cfg
  RNB_FUSE=$1 RNB_FUSE_NEXT=$2 timeout 60 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/prof${tag}_$1$2.json > gpurun_out/bench${tag}_$1$2.json 2> gpurun_out/bench${tag}_$1$2.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench${tag}_$1$2.json").read().strip().splitlines()[-1])
L=json.load(open("gpurun_out/prof${tag}_$1$2.json"))["launches"]
print("fuse $1 next $2:", round(d["ms_per_step"],4), round(d["value"]), round(d["roofline"]["frac"],4), d["gpu_launches"], "|", " ".join(f"{l['ms']*1000:.0f}" for l in L[:12]))
P
done
