"""A/B of the host-input serving loop (rnb_model_submit_host / wait_host, two slots): plain FP32 copies, every image
rounded to BF16 on the host cores, fixed fractions, the model's own choice, and uint8 input — on one box, interleaved.

    python tools/e2e_ab.py [arch] [batch] [dtype] [steps]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from resnet_c_b200 import engine, weights  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 40

model = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, False), dtype=dtype, max_batch=B)
xs = [weights.synthetic_images(B, seed=s).pin_memory() for s in (1, 2)]
xu = [weights.synthetic_images_u8(B, seed=s).pin_memory() for s in (3, 4)]
lh = [torch.empty(B, model.num_classes).pin_memory() for _ in range(2)]
th = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
xd = xs[0].cuda()
model.forward(xd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    model.forward(xd)
e0.record()
for _ in range(20):
    model.forward(xd)
e1.record()
torch.cuda.synchronize()
print(f"{arch} B={B} {dtype}: device-resident step {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)


def loop(u8):
    sub = model.submit_host_u8 if u8 else model.submit_host
    x = xu if u8 else xs
    for i in range(2):
        sub(i, x[i], lh[i], th[i])
    for i in range(2):
        model.wait_host(i)
    t0 = time.perf_counter()
    sub(0, x[0], lh[0], th[0])
    for i in range(1, steps):
        sub(i & 1, x[i & 1], lh[i & 1], th[i & 1])
        model.wait_host((i - 1) & 1)
    model.wait_host((steps - 1) & 1)
    return (time.perf_counter() - t0) / steps * 1e3


model.submit_host(0, xs[0], lh[0], th[0])   # the model's own decision (made at its first host call)
model.wait_host(0)
auto = model.host_pack()
print("auto decision:", auto, flush=True)
forms = [("plain FP32 copies", 0.0), ("25 % on the host", 0.25), ("50 %", 0.5), ("75 %", 0.75), ("all on the host", 1.0),
         (f"auto ({auto['fraction']:.2f})", auto["fraction"])]
for rnd in range(2):
    for name, f in forms:
        model.set_host_pack_fraction(f)
        ms = loop(False)
        print(f"round {rnd}: {name:22s} {ms:.3f} ms per step  {B / ms:.1f} k images/s", flush=True)
    ms = loop(True)
    print(f"round {rnd}: {'uint8 input':22s} {ms:.3f} ms per step  {B / ms:.1f} k images/s", flush=True)
