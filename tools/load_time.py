"""GPU box: start-up time from the reference's weight directory (one raw file per state_dict key, BN folded and
repacked at load) vs from the pre-packed blob (rnb_model_create_packed). Usage: python tools/load_time.py [arch]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
wdir = weights.cached_weights_dir(arch, 0, True)
nfiles = len(list(Path(wdir).iterdir()))
torch.cuda.init()
m = engine.ResNet(arch, wdir, dtype="bf16", max_batch=8)   # warm-up (context, module load)
blob = Path(f"/tmp/{arch}_bf16.rnbw")
m.save_packed(blob)
m.close()
for label, make in (("weights directory", lambda: engine.ResNet(arch, wdir, dtype="bf16", max_batch=8)),
                    ("packed blob", lambda: engine.ResNet.from_packed(blob, max_batch=8))):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mm = make()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        mm.close()
    print(f"{arch} bf16 from {label}: {min(ts) * 1e3:.1f} ms (best of 5)"
          + (f"  [{nfiles} files]" if label.startswith("weights") else f"  [{blob.stat().st_size / 1e6:.1f} MB, 1 file]"))
