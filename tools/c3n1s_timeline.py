"""GPU: in-kernel timeline of bneck_c3n1s_kernel (layer3 conv3 + shortcut + ReLU + next conv1).

Build the instrumented library first (the product library compiles the probes away):
  cd resnet_c_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DRNB_TIMELINE \
      -shared -o ../../build/librnb_tl.so api.cu model.cu conv_plan.cu tensormap.cu ops_f32.cu layout.cu stem.cu \
      stem_tc.cu stem_tc_split.cu tail.cu
Then: python tools/c3n1s_timeline.py  -> merged timeline (clocks relative to the first event) of the SECOND tile of CTA 0:
role E = epilogue warp 4, C3 = conv3 issuer, C1 = conv1' issuer, ST = store warp."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["RNB_AUTOTUNE"] = "0"
os.environ["RNB_NO_GRAPH"] = "1"
import torch  # noqa: E402
from resnet_c_b200 import _lib  # noqa: E402

_lib.LIB_PATH = ROOT / "build" / "librnb_tl.so"
from resnet_c_b200 import engine, weights  # noqa: E402

m = engine.ResNet("resnet50", weights.cached_weights_dir("resnet50", 0, True), dtype="bf16", max_batch=256)
x = weights.synthetic_images(256).cuda()
for _ in range(3):
    m.forward(x)
torch.cuda.synchronize()
lib = _lib.lib()
KT = 160
buf = (C.c_longlong * (4 * KT))()
cnt = (C.c_int * 4)()
lib.rnb_debug_read_timeline.restype = C.c_int
assert lib.rnb_debug_read_timeline(buf, cnt) == KT
ROLE = {0: "E ", 1: "C3", 2: "C1", 3: "ST"}


def name(role, tag):
    if role == 0:
        if tag < 10:
            return {1: "wait d2_full[0]", 2: "got d2_full[0]", 3: "wait d2_full[1]", 4: "got d2_full[1]", 5: "wait d3_full",
                    6: "got d3_full", 7: "tile done"}[tag]
        k, it = divmod(tag - 20, 30)
        return ["wait box_ready", "got box_ready", "tmem loaded", "box published"][k] + f" item {it}"
    if role == 1:
        if tag < 10:
            return {1: "wait a_full", 2: "got a_full"}[tag]
        k, c = divmod(tag - 10, 10)
        return ["wait d2_empty", "got d2_empty", "got w3_full"][k] + f" chunk {c}"
    if role == 2:
        if tag < 10:
            return {1: "wait d3_empty", 2: "got d3_empty"}[tag]
        return ("got w1_full" if tag < 30 else "got cx_full -> issue") + f" box {(tag - 10) % 20}"
    if tag >= 90:
        return f"recycled item {tag - 90}"
    if tag >= 65:
        return f"store issued + prev read out, item {tag - 65}"
    if tag >= 40:
        return f"got c_full item {tag - 40}"
    return f"wait c_full item {tag - 10}"


ev = []
for r in range(4):
    for i in range(cnt[r]):
        v = buf[r * KT + i]
        ev.append((v >> 8, r, v & 255))
ev.sort()
t0 = ev[0][0]
print("events per role:", list(cnt))
last = {}
for t, r, tag in ev:
    dt = t - last.get(r, t)
    last[r] = t
    print(f"{t - t0:8d}  {ROLE[r]}  +{dt:6d}  {name(r, tag)}")
print("tile span (clk):", ev[-1][0] - t0)
