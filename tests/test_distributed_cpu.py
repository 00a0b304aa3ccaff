"""world_size-2 gloo test (CPU) of the multi-GPU host logic: image sharding + confusion-matrix all-gather
must reproduce the single-process confusion matrix / mIoU bit for bit."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _conf_cpu(pred, gt, ncls, ignore=255):
    m = gt != ignore
    return torch.bincount(gt[m].long() * ncls + pred[m].long(), minlength=ncls * ncls).view(ncls, ncls)


def _data(n=6, ncls=25):
    g = torch.Generator().manual_seed(4)
    pred = torch.randint(0, ncls, (n, 64, 48), generator=g, dtype=torch.uint8)
    gt = torch.randint(0, ncls, (n, 64, 48), generator=g, dtype=torch.uint8)
    gt[torch.rand(gt.shape, generator=g) < 0.01] = 255
    return pred, gt


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmsam_b200  # noqa
    from mmsam_b200.evalmetrics import gather_confusion, shard_indices
    pred, gt = _data()
    idx = shard_indices(pred.shape[0], rank, world)
    conf = _conf_cpu(pred[idx], gt[idx], 25)
    total = gather_confusion(conf)
    q.put((rank, idx, total))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_confusion_allgather_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    import mmsam_b200  # noqa
    from mmsam_b200.evalmetrics import areas_from_confusion, metrics_from_confusion
    pred, gt = _data()
    want = _conf_cpu(pred, gt, 25)
    seen = sorted(i for _, idx, _ in res for i in idx)
    assert seen == list(range(pred.shape[0]))            # every image exactly once
    for _, _, total in res:
        assert torch.equal(total, want)                   # integer counts: bit-identical on every rank
    # the areas equal the reference's histc-based intersect_and_union (metrics_micro.py:74-86)
    m = gt != 255
    p, l = pred[m].float(), gt[m].float()
    inter = torch.histc(p[p == l], bins=25, min=0, max=24)
    ap, al = torch.histc(p, bins=25, min=0, max=24), torch.histc(l, bins=25, min=0, max=24)
    i2, u2, p2, l2 = areas_from_confusion(want)
    assert torch.equal(i2.float(), inter) and torch.equal(p2.float(), ap) and torch.equal(l2.float(), al)
    assert torch.equal(u2.float(), ap + al - inter)
    mt = metrics_from_confusion(want)
    assert abs(mt["mIoU"] - torch.nanmean(inter / (ap + al - inter)).item()) < 1e-6   # fp32 vs fp64 division


def test_deliver_keys_and_bucket_metrics_host_logic():
    """Per-condition bucketing (datasets/DELIVER.py:261-615): path -> (weather, case) keys; bucket matrices add up to
    the global one; metrics per bucket follow total_area_to_metrics."""
    sys.path.insert(0, ROOT)
    import mmsam_b200  # noqa
    from mmsam_b200 import evalmetrics as em
    assert em.deliver_keys("data/DELIVER/img/cloud/test/MAP_10_point102/045050_rgb_front.png") == ("cloud", None)
    assert em.deliver_keys("data/DELIVER/lists/test_motionblur/img/night/test/MAP_1/000001_rgb_front.png") == ("night", "motionblur")
    pred, gt = _data(n=6)
    keys = [("cloud", "ordinary"), ("fog",), ("cloud",), (), ("sun", "lidarjitter"), ("cloud", "unknown-key")]
    bc = em.BucketedConfusion(25, em.DELIVER_WEATHERS + em.DELIVER_CASES, "cpu")
    for i in range(pred.shape[0]):                       # CPU stand-in for the device kernel: same counts
        c = _conf_cpu(pred[i], gt[i], 25)
        for k in ["global"] + [k for k in keys[i] if k in bc.index]:
            bc.conf[bc.index[k]] += c
    assert torch.equal(bc.conf[0], _conf_cpu(pred, gt, 25))
    assert torch.equal(bc.conf[bc.index["cloud"]], _conf_cpu(pred[[0, 2, 5]], gt[[0, 2, 5]], 25))
    assert int(bc.conf[bc.index["rain"]].sum()) == 0
    m = bc.metrics()
    assert set(m) == set(bc.keys) and abs(m["global"]["mIoU"] - em.metrics_from_confusion(bc.conf[0])["mIoU"]) < 1e-12
    assert torch.equal(bc.gather(), bc.conf)             # no process group: identity
