"""GPU check: the TF32 halo-pair 3x3 kernel (csrc/conv3x3_halo2.cuh) against the generic im2col kernel
(RNB_NO_HALO=1): ResNet-18 TF32 block outputs and logits must be bit-identical."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402


def run(arch, batch, env):
    from resnet_c_b200 import engine, weights
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    os.environ["RNB_KEEP_ACTIVATIONS"] = "1"
    os.environ["RNB_AUTOTUNE"] = "0"
    try:
        m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype="tf32", max_batch=batch)
        logits, top1 = m.forward(weights.synthetic_images(batch).cuda())
        torch.cuda.synchronize()
        acts = {n: m.activation(n).clone() for n in ("maxpool", "layer1.0", "layer1.1", "layer2.0")}
        m.close()
        return logits.clone(), top1.clone(), acts
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


ok = True
for arch, b in (("resnet18", 3), ("resnet18", 37), ("resnet34", 2)):
    base = run(arch, b, {"RNB_NO_HALO": "1"})
    got = run(arch, b, {"RNB_NO_HALO": ""})
    os.environ.pop("RNB_NO_HALO", None)
    got = run(arch, b, {})
    line = [f"{arch} B={b}"]
    for n in base[2]:
        d = (base[2][n] - got[2][n]).abs().max().item()
        line.append(f"{n} max|d|={d:.3g}")
        ok = ok and d == 0.0 and not bool(torch.isnan(got[2][n]).any())
    ok = ok and torch.equal(base[0], got[0]) and torch.equal(base[1], got[1])
    line.append(f"logits equal={torch.equal(base[0], got[0])}")
    print(" | ".join(line), flush=True)
print("HALO2_CHECK", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
