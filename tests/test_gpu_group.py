"""-m gpu: rnb_group_* — several replicas driven by ONE process through the C ABI.

The reference has one device and B = 1 (cuda/inference/main.cu:230), so the oracle for the sharded path is the
single-replica path itself: the gathered logits / top-1 of a sharded batch must equal, BIT FOR BIT, one model
forwarding the whole batch (every image is independent, the kernels are deterministic and batch-size independent —
tests/test_gpu_model.py checks that independence against the oracle)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _single(arch, wdir, x, dtype="bf16"):
    from resnet_c_b200 import engine
    m = engine.ResNet(arch, wdir, dtype=dtype, max_batch=x.shape[0])
    logits, top1 = m.forward(x.cuda())
    torch.cuda.synchronize()
    out = logits.cpu(), top1.cpu()
    m.close()
    return out


def _shards(group, x):
    out = []
    for r, dev in enumerate(group.devices):
        lo, n = group.shard(x.shape[0], r)
        out.append(x[lo:lo + n].to(f"cuda:{dev}"))
    return out


@pytest.mark.parametrize("batch", [6, 5, 1])   # even, ragged, fewer images than replicas
def test_two_replicas_on_one_gpu_equal_one_model(batch):
    from resnet_c_b200 import engine, weights
    wdir = weights.cached_weights_dir("resnet18", 0, True)
    x = weights.synthetic_images(batch)
    want_l, want_t = _single("resnet18", wdir, x)
    g = engine.ResNetGroup("resnet18", wdir, [0, 0], max_batch_per_device=4)
    assert len(g) == 2
    assert [g.shard(batch, r) for r in range(2)] == [(0, (batch + 1) // 2), ((batch + 1) // 2, batch // 2)]
    logits, top1 = g.forward(_shards(g, x))
    g.synchronize()
    assert torch.equal(logits.cpu(), want_l) and torch.equal(top1.cpu(), want_t)
    hl, ht = g.forward_host(x.pin_memory())
    assert torch.equal(hl, want_l) and torch.equal(ht, want_t)
    g.close()


def test_group_u8_input_and_slots():
    from resnet_c_b200 import engine, weights
    wdir = weights.cached_weights_dir("resnet50", 0, True)
    xu = weights.synthetic_images_u8(6, seed=3)
    m = engine.ResNet("resnet50", wdir, max_batch=6)
    want_l, want_t = m.forward_u8(xu.cuda())
    torch.cuda.synchronize()
    g = engine.ResNetGroup("resnet50", wdir, [0, 0], max_batch_per_device=3)
    logits, top1 = g.forward(_shards(g, xu))
    g.synchronize()
    assert torch.equal(logits, want_l) and torch.equal(top1, want_t)
    bufs = [(torch.empty(6, g.num_classes).pin_memory(), torch.empty(6, dtype=torch.int32).pin_memory()) for _ in range(2)]
    xh = xu.pin_memory()
    for s in range(2):
        g.submit_host(s, xh, *bufs[s])
    for s in range(2):
        g.wait_host(s)
        assert torch.equal(bufs[s][0], want_l.cpu()) and torch.equal(bufs[s][1], want_t.cpu())
    g.close()
    m.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("gather", ["direct", "copy"])
def test_two_gpus_gathered_logits_are_bit_exact(gather, monkeypatch):
    """Replica 1 lives on GPU 1 and stores its rows straight into GPU 0's buffers (peer-mapped TMA store of the FC
    kernel + arg-max stores), or through cudaMemcpyPeerAsync with RNB_GROUP_GATHER=copy."""
    from resnet_c_b200 import engine, weights
    if gather == "copy":
        monkeypatch.setenv("RNB_GROUP_GATHER", "copy")
    wdir = weights.cached_weights_dir("resnet50", 0, True)
    x = weights.synthetic_images(9)
    want_l, want_t = _single("resnet50", wdir, x)
    g = engine.ResNetGroup("resnet50", wdir, [0, 1], max_batch_per_device=5)
    assert g.direct_stores(1) == (gather == "direct")
    for _ in range(3):   # graph replay included
        logits, top1 = g.forward(_shards(g, x))
    g.synchronize()
    assert logits.device.index == 0
    assert torch.equal(logits.cpu(), want_l) and torch.equal(top1.cpu(), want_t)
    hl, ht = g.forward_host(x.pin_memory())
    assert torch.equal(hl, want_l) and torch.equal(ht, want_t)
    g.close()
