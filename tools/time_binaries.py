"""BASELINE.md section 5 context timing: the reference's own CUDA program and the two drop-in programs, ResNet-152, the
reference test image, batch 1, whole-process wall clock (weight load + forward + arg-max; the reference's main.cu has no
timer of its own and is run UNMODIFIED, so the process is the unit). Run on a GPU box:

    python tools/time_binaries.py > gpurun_out/time_binaries.json
"""
import json
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from resnet_c_b200 import weights  # noqa: E402

BINS = {
    "reference cuda_inference_out (unmodified main.cu + reference ops.cu/nn.cu)": [ROOT / "oracle/_ref/cuda_inference_out"],
    "ref_main_dropin (unmodified main.cu on this repo's cuda/*.cuh + librnb.so, FP32 per-op kernels)": [ROOT / "build/ref_main_dropin"],
    "resnet_infer resnet152 bf16 (whole-model object, tcgen05 path)": [ROOT / "build/resnet_infer", "resnet152", "bf16", "1"],
    "resnet_infer resnet152 tf32": [ROOT / "build/resnet_infer", "resnet152", "tf32", "1"],
}


def main():
    d = Path(tempfile.mkdtemp(prefix="r152_"))
    (d / "weights_bin").symlink_to(weights.cached_weights_dir("resnet152", 0))
    (d / "test_bins").mkdir()
    (d / "test_bins" / "ILSVRC2012_val_00004749.bin").symlink_to(ROOT / "tests/golden/ILSVRC2012_val_00004749.bin")
    out = {}
    for name, cmd in BINS.items():
        if not Path(cmd[0]).exists():
            out[name] = {"skipped": "binary not built"}
            continue
        runs = []
        for _ in range(2):  # second run: files in the page cache
            t0 = time.perf_counter()
            r = subprocess.run([str(c) for c in cmd], cwd=d, capture_output=True, text=True, timeout=1800)
            runs.append(time.perf_counter() - t0)
        m = re.search(r"Finished \(([\d.]+) ms", r.stdout)
        out[name] = {"wall_s": [round(t, 3) for t in runs], "rc": r.returncode,
                     "top1": [int(x) for x in re.findall(r"max index is (\d+)", r.stdout)],
                     "forward_ms_reported": float(m.group(1)) if m else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
