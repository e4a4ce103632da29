"""Developer check (GPU): whole-model parity of the engine against the CPU oracle, layer by layer.

  python tools/e2e_check.py [arch] [dtype] [batch] [--rbn] [--taps]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

from oracle import torch_model
from resnet_c_b200 import engine, weights


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    arch = args[0] if len(args) > 0 else "resnet18"
    dtype = args[1] if len(args) > 1 else "bf16"
    batch = int(args[2]) if len(args) > 2 else 2
    rbn = "--rbn" in sys.argv
    want_taps = "--taps" in sys.argv
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = weights.make_state_dict(arch, 0, randomize_bn=rbn)
    wdir = weights.cached_weights_dir(arch, 0, rbn)
    x = weights.synthetic_images(batch)
    taps = {} if want_taps else None
    t0 = time.time()
    ref = torch_model.run(arch, sd, x, taps=taps)
    print(f"oracle fp32 CPU: {time.time() - t0:.2f}s")
    model = engine.ResNet(arch, wdir, dtype=dtype, max_batch=batch)
    xd = x.cuda()
    logits, top1 = model.forward(xd)
    torch.cuda.synchronize()
    got = logits.cpu()
    rel = ((got - ref).abs().amax(1) / ref.abs().amax(1)).max().item()
    print(f"{arch} {dtype} B={batch} rbn={rbn}: logits rel err (max|d|/max|y| per image) = {rel:.3e}")
    print("top1 engine", top1.cpu().tolist()[:8], "oracle", ref.argmax(1).tolist()[:8],
          "match", bool((top1.cpu() == ref.argmax(1).to(torch.int32)).all()))
    if want_taps:
        for name, t in taps.items():
            a = model.activation(name).cpu().reshape(t.shape)
            e = (a - t).abs().max().item() / max(t.abs().max().item(), 1e-30)
            print(f"  {name:12s} rel err {e:.3e}  (max|ref| {t.abs().max().item():.4g})")
    # host path
    lh, th = model.forward_host(x.pin_memory())
    print("forward_host max |diff| vs device path:", (lh - got).abs().max().item(),
          "top1 equal", bool((th == top1.cpu()).all()))


if __name__ == "__main__":
    main()
