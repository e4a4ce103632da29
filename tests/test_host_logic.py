"""Host-side logic that needs no GPU: on-disk formats (save_weights.py / convert_imgs_to_bin.py),
shard arithmetic, and the N>1 gather path run as a real world_size-2 gloo job on CPU."""
import os
import struct
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT
from resnet_c_b200 import dist as rdist
from resnet_c_b200 import weights


def test_weight_files_match_the_reference_writer(tmp_path):
    """save_weights.py:9-12 writes each element with struct.pack('f', x.item()); ours must be
    byte-identical, including the 4-byte float it makes of the int64 num_batches_tracked."""
    sd = {
        "conv1.weight": torch.randn(4, 3, 2, 2),
        "bn1.running_var": torch.rand(4) + 0.5,
        "bn1.num_batches_tracked": torch.tensor(7, dtype=torch.int64),
    }
    n = weights.save_weights_dir(sd, tmp_path)
    assert n == 3
    for name, t in sd.items():
        expected = b"".join(struct.pack("f", x.item()) for x in t.data.flatten())
        assert (tmp_path / name).read_bytes() == expected
    assert (tmp_path / "bn1.num_batches_tracked").stat().st_size == 4


def test_weight_dir_round_trip_and_file_inventory(tmp_path):
    sd = weights.make_state_dict("resnet18", 0, randomize_bn=True)
    n = weights.save_weights_dir(sd, tmp_path)
    assert n == 122  # ResNet-18 = 122 files
    back = weights.load_weights_dir("resnet18", tmp_path)
    for k, v in sd.items():
        assert torch.equal(back[k], v), k
    # names the reference's loaders ask for (nn.cuh:21,58-61,113-117; main.cu:59-75)
    for f in ("conv1.weight", "bn1.running_mean", "layer2.0.downsample.0.weight",
              "layer2.0.downsample.1.running_var", "layer4.1.conv2.weight", "fc.weight", "fc.bias"):
        assert (tmp_path / f).exists()


def test_image_bin_round_trip(tmp_path, jpeg_tensor):
    assert jpeg_tensor.shape == (1, 3, 224, 224)
    p = tmp_path / "img.bin"
    weights.save_image_bin(jpeg_tensor, p)
    assert p.stat().st_size == 3 * 224 * 224 * 4  # convert_imgs_to_bin.py:20-23
    assert torch.equal(weights.load_image_bin(p), jpeg_tensor)
    # statistics recorded by the survey for this fixture
    assert float(jpeg_tensor.min()) == pytest.approx(-1.98, abs=0.01)
    assert float(jpeg_tensor.max()) == pytest.approx(2.45, abs=0.01)


@pytest.mark.skipif(not (REFERENCE / "test_imgs").exists(), reason="reference tree not mounted")
def test_preprocess_reproduces_committed_image(jpeg_tensor):
    live = weights.preprocess_jpeg(REFERENCE / "test_imgs" / "ILSVRC2012_val_00004749.jpeg")
    assert torch.equal(live, jpeg_tensor)


def test_shard_bounds_cover_the_batch_exactly():
    for total in (0, 1, 7, 256, 2048, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [rdist.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b - a >= d - c >= 0
    with pytest.raises(ValueError):
        rdist.shard_bounds(8, 2, 2)


_WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from resnet_c_b200 import dist as rdist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    total, classes = {total}, 10
    g = torch.Generator().manual_seed(5)
    full = torch.randn(total, classes, generator=g)
    lo, hi = rdist.shard_bounds(total, world, rank)
    logits = full[lo:hi].clone()
    top1 = logits.argmax(1).to(torch.int32) if hi > lo else torch.empty(0, dtype=torch.int32)
    all_l, all_t = rdist.gather_results(logits, top1, total)
    assert torch.equal(all_l, full), "gathered logits differ"
    assert torch.equal(all_t, full.argmax(1).to(torch.int32)), "gathered top1 differ"
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


@pytest.mark.parametrize("total", [8, 7, 1])
def test_gather_world_size_2_gloo(tmp_path, total):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT), total=total))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() + total) % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


_ANY_WORKER = textwrap.dedent("""
    import sys, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from resnet_c_b200 import dist as rdist
    assert rdist.any_rank(True) is True and rdist.any_rank(False) is False     # no process group: this rank alone
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    # rank 1 alone decides "yes" (a replica that chose host packing): both ranks must take the branch, and the
    # collectives inside it must pair up
    for flags in ((False, True), (True, False), (False, False), (True, True)):
        got = rdist.any_rank(flags[rank])
        assert got == any(flags), (flags, rank, got)
        if got:
            dist.barrier()
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_rank_local_decisions_are_agreed_before_collectives_gloo(tmp_path):
    """bench.py times the plain-copy serving loop (barriers inside) when the model chose host packing; every rank
    decides on its own timings, so the branch is taken on all ranks or none (an 8-GPU run hung on exactly this)."""
    script = tmp_path / "worker.py"
    script.write_text(_ANY_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() + 977) % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
    # and bench.py goes through it
    assert "rdist.any_rank(packed" in (ROOT / "bench.py").read_text()


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` prints ONE JSON line with the keys the driver reads."""
    import json
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-sample", "2", "--arch", "resnet18"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0
    assert d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_c_abi_shard_rule_equals_python_rule():
    """rnb_shard_bounds (what rnb_group_* shards by, csrc/group.cu) and dist.shard_bounds (what the torchrun path
    shards by) are the same rule — a batch split by one host can be gathered by the other. No GPU involved."""
    import ctypes as C

    from resnet_c_b200 import _lib, dist
    lib = _lib.lib()
    for total in (0, 1, 7, 256, 1000, 2048):
        for world in (1, 2, 3, 8):
            covered = 0
            for rank in range(world):
                first, count = C.c_int(), C.c_int()
                assert lib.rnb_shard_bounds(total, world, rank, C.byref(first), C.byref(count)) == 0
                lo, hi = dist.shard_bounds(total, world, rank)
                assert (first.value, first.value + count.value) == (lo, hi)
                covered += count.value
            assert covered == total
    first, count = C.c_int(), C.c_int()
    assert lib.rnb_shard_bounds(10, 0, 0, C.byref(first), C.byref(count)) != 0
    assert b"bad argument" in lib.rnb_last_error()


def _bf16_bits_torch(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


@pytest.mark.parametrize("n, offset", [(0, 0), (1, 0), (15, 0), (16, 0), (17, 3), (32768, 0), (32769, 1), (3 * 224 * 224 * 5 + 7, 0)])
def test_host_bf16_rounding_is_round_to_nearest_even(n, offset):
    """rnb_f32_to_bf16_host (the host cores' half of the packed host paths, csrc/host_pack.cpp) against torch's
    FP32 -> BF16 cast, bit for bit: random values of every magnitude, ties, denormals, infinities; ragged sizes and
    a destination that is not 32-byte aligned (vector body + scalar tail, streaming and plain stores)."""
    from resnet_c_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(1234 + n)
    bits = torch.randint(-2**31, 2**31 - 1, (n,), generator=g, dtype=torch.int64).to(torch.int32)
    x = bits.view(torch.float32).clone()
    if n >= 16:
        special = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1.0, 1.00390625, 1.01171875, 3.3895314e38,
                                1e-40, -1e-40, 1.1754944e-38, 65504.0, 0.5 + 2**-9, 0.5 + 3 * 2**-9, -2.5, 1e-45],
                               dtype=torch.float32)
        x[:16] = special
    finite_or_inf = ~torch.isnan(x)
    out = np.zeros(n + 64, dtype=np.uint16)
    dst = out[offset:offset + n]
    assert lib.rnb_f32_to_bf16_host(x.data_ptr(), dst.ctypes.data, n) == 0
    want = _bf16_bits_torch(x)
    m = finite_or_inf.numpy()
    np.testing.assert_array_equal(dst[m], want[m])
    # NaN: the GPU's cvt.rn.bf16.f32 produces the canonical 0x7FFF, and so does the host
    assert (dst[~m] == 0x7FFF).all()
    assert (out[:offset] == 0).all() and (out[offset + n:] == 0).all()   # nothing written outside [0, n)
    assert lib.rnb_host_pack_threads() >= 1


def test_host_bf16_rounding_thread_count_follows_the_environment():
    code = ("from resnet_c_b200 import _lib; import numpy as np; l = _lib.lib(); print(l.rnb_host_pack_threads());"
            "x = np.linspace(-3, 3, 200001, dtype=np.float32); o = np.zeros_like(x, dtype=np.uint16);"
            "assert l.rnb_f32_to_bf16_host(x.ctypes.data, o.ctypes.data, x.size) == 0;"
            "print(int(o.astype(np.uint64).sum()))")
    outs = []
    # (default = AVX2 where the CPU has it; the other two runs take the AVX-512 form, if the CPU has it, and the scalar one)
    for threads, isa in (("1", ""), ("3", "avx512"), ("2", "scalar")):
        env = dict(os.environ, RNB_HOST_THREADS=threads, RNB_HOST_ISA=isa, PYTHONPATH=str(ROOT))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        lines = r.stdout.split()
        assert lines[0] == threads
        outs.append(lines[1])
    assert outs[0] == outs[1] == outs[2]   # neither the partitioning nor the instruction set changes the result


def test_host_pack_pool_is_race_free_under_thread_sanitizer(tmp_path):
    """csrc/host_pack.cpp under -fsanitize=thread: three caller threads share the process-wide pool; pieces must be
    announced once, in order and complete, results equal the scalar conversion, and TSAN must stay silent."""
    exe = tmp_path / "host_pack_stress"
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-o", str(exe),
                            str(ROOT / "tests" / "host_pack_stress.cpp"),
                            str(ROOT / "resnet_c_b200" / "csrc" / "host_pack.cpp")], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("no thread sanitizer runtime in this toolchain")
    assert build.returncode == 0, build.stderr
    for threads in ("1", "6"):
        r = subprocess.run([str(exe)], env=dict(os.environ, RNB_HOST_THREADS=threads), capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "failures 0" in r.stdout and f"threads {threads} " in r.stdout
        assert "ThreadSanitizer" not in r.stderr


def test_host_bf16_rounding_survives_a_fork():
    """The pool's threads do not exist in a forked child: the conversion must still complete there (on the calling
    thread), not wait for workers that are gone."""
    code = textwrap.dedent("""
        import os, sys, numpy as np
        from resnet_c_b200 import _lib
        l = _lib.lib()
        x = np.linspace(-5, 5, 300001, dtype=np.float32)
        want = np.zeros(x.size, dtype=np.uint16)
        assert l.rnb_f32_to_bf16_host(x.ctypes.data, want.ctypes.data, x.size) == 0      # starts the pool
        pid = os.fork()
        if pid == 0:
            got = np.zeros(x.size, dtype=np.uint16)
            ok = l.rnb_f32_to_bf16_host(x.ctypes.data, got.ctypes.data, x.size) == 0 and (got == want).all()
            os._exit(0 if ok else 3)
        _, status = os.waitpid(pid, 0)
        print("child", os.WEXITSTATUS(status) if os.WIFEXITED(status) else -1)
    """)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RNB_HOST_THREADS="4", PYTHONPATH=str(ROOT)),
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "child 0" in r.stdout


def test_input_reference_sub_batch_arithmetic(tmp_path):
    """InRef (csrc/model.h): offsets into FP32 / uint8 / mixed BF16 + FP32 batches as the chunk loop and the two-lane
    split take them — host-only program built with nvcc (tests/inref_check.cu), no GPU involved."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    exe = tmp_path / "inref_check"
    b = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe),
                        str(ROOT / "tests" / "inref_check.cu")], capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "inref ok" in r.stdout, r.stdout + r.stderr


def test_host_pack_split_rule():
    """rnb_host_pack_split (csrc/host_pack.cpp): the share of a batch the host cores round to BF16, from measured
    rates — checked on the figures the B200 boxes printed (profiles/e2e_hostpack_ab_r2.txt, hostpack_8gpu_r2.txt)."""
    import ctypes as C
    from resnet_c_b200 import _lib
    lib = _lib.lib()

    def split(conv, f32, bf16, batch=256):
        n = C.c_int(-1)
        f = lib.rnb_host_pack_split(conv, f32, bf16, batch, C.byref(n))
        return f, n.value

    f, n = split(76.1, 53.8, 52.3)                 # one GPU, 16 cores, pinned input: 73 % -> 192 of 256 images
    assert abs(f - 0.73) < 0.01 and n == 192
    f, n = split(74.8, 5.5, 52.3)                  # pageable input (the plain copy crawls): everything through the cores
    assert f == 1.0 and n == 256
    for conv, f32, bf16 in [(12.8, 51.2, 38.7), (9.9, 39.8, 34.3), (10.9, 24.5, 20.2)]:   # 8 ranks share 32 cores
        assert split(conv, f32, bf16) == (0.0, 0)
    f, n = split(12.0, 24.3, 28.8)                 # ... the one rank that saw a slow link at that moment
    assert abs(f - 0.32) < 0.01 and n == 80
    assert split(0.0, 50.0, 50.0) == (0.0, 0) and split(50.0, 0.0, 50.0) == (0.0, 0)
    # whole 16-image pieces; short remainders join the other side; small batches are not split
    assert [split(76.1, 53.8, 52.3, b)[1] for b in (37, 40, 70, 128, 31, 8, 1)] == [37, 40, 48, 96, 31, 8, 1]
    assert [split(20.0, 40.0, 40.0, b)[1] for b in (256, 31, 8)] == [80, 0, 0]    # f = 0.33: the larger side is the link
    # by construction both sides finish together: core time = link time at the returned fraction
    f, _ = split(40.0, 50.0, 50.0)
    core, link = f / (40.0 * 0.8), f * 0.5 / 50.0 + (1 - f) / 50.0
    assert 0.2 < f < 0.93 and abs(core - link) < 1e-12
