"""GPU check of the fused layer1 Bottleneck tail (csrc/bneck_l1.cuh) against the layer-by-layer path.

RNB_FUSE=1 (conv2 + conv3 + residual, optional next conv1) rounds at the same points as the unfused
kernels, so block outputs and logits must be BIT-IDENTICAL; RNB_FUSE=2 (downsample folded into the
conv3 accumulator) skips one BF16 rounding of the shortcut, so block 0 is compared with a tolerance.
Usage: python tools/fuse_check.py [arch] [batch ...]
"""
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
    try:
        m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype="bf16", max_batch=batch)
        x = weights.synthetic_images(batch).cuda()
        logits, top1 = m.forward(x)
        torch.cuda.synchronize()
        acts = {n: m.activation(n).clone() for n in ("maxpool", "layer1.2", "layer2.0", "layer2.1", "layer2.2", "layer2.3", "layer3.0", "layer3.1", "layer3.3", "layer3.5", "layer4.0")}
        out = (logits.clone(), top1.clone(), acts, m.launches_per_forward(batch))
        m.close()
        return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
    batches = [int(a) for a in sys.argv[2:]] or [2, 37]
    ok = True
    for b in batches:
        base = run(arch, b, {"RNB_FUSE": "0"})
        for tag, env, exact in [
            ("fuse=1 next=0", {"RNB_FUSE": "1", "RNB_FUSE_NEXT": "0"}, True),
            ("fuse=1 next=1", {"RNB_FUSE": "1", "RNB_FUSE_NEXT": "1"}, True),
            ("fuse=2 next=0", {"RNB_FUSE": "2", "RNB_FUSE_NEXT": "0"}, False),
            ("fuse=2 next=1", {"RNB_FUSE": "2", "RNB_FUSE_NEXT": "1"}, False),
        ]:
            got = run(arch, b, env)
            line = [f"{arch} B={b} {tag}: launches {base[3]} -> {got[3]}"]
            for n in base[2]:
                a, c = base[2][n], got[2][n]
                d = (a - c).abs().max().item()
                rel = d / a.abs().max().item()
                nan = bool(torch.isnan(c).any())
                line.append(f"{n} max|d|={d:.3g} rel={rel:.2e}{' NaN!' if nan else ''}")
                if nan or (exact and d != 0.0) or rel > 3e-2:
                    ok = False
            dl = (base[0] - got[0]).abs().max().item() / base[0].abs().max().item()
            same_top1 = bool((base[1] == got[1]).all())
            line.append(f"logits rel={dl:.2e} top1_equal={same_top1}")
            if (exact and dl != 0.0) or dl > 1e-2:
                ok = False
            print(" | ".join(line), flush=True)
    print("FUSE_CHECK", "OK" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
