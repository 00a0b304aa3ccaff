"""Evaluation glue that follows the path: per-rank confusion matrices (device-side, kernels.confusion),
one all-gather across ranks, then the reference's metrics.

intersect / union / pred / label areas of metrics_micro.intersect_and_union
(mmseg_custom/apis/evaluation/metrics_micro.py:26-86) are the diagonal / row+col sums of the confusion matrix;
metrics follow total_area_to_metrics (:451-526). Integer counts, so the multi-GPU result is bit-identical
to the single-GPU one."""
import torch


def shard_indices(n_items, rank, world):
    """Round-robin image sharding (image i -> rank i mod world), like a DistributedSampler without shuffle."""
    return list(range(rank, n_items, world))


def gather_confusion(conf, group=None):
    """Sum the per-rank [C,C] int64 confusion matrices on every rank (the path's only collective)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return conf.clone()
    parts = [torch.zeros_like(conf) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, conf.contiguous(), group=group)
    return torch.stack(parts).sum(0)


def areas_from_confusion(conf):
    """-> (area_intersect, area_union, area_pred_label, area_label), each [C] int64 (rows = gt, cols = pred)."""
    inter = conf.diagonal()
    label = conf.sum(1)
    pred = conf.sum(0)
    return inter, pred + label - inter, pred, label


def metrics_from_confusion(conf):
    """aAcc / IoU / Acc / mIoU as in total_area_to_metrics(metrics=['mIoU']) + nanmean over classes."""
    inter, union, pred, label = (t.double() for t in areas_from_confusion(conf.cpu()))
    iou = inter / union
    acc = inter / label
    return dict(aAcc=(inter.sum() / label.sum()).item(), IoU=iou, Acc=acc, mIoU=torch.nanmean(iou).item(),
                mAcc=torch.nanmean(acc).item())


DELIVER_WEATHERS = ("cloud", "fog", "night", "rain", "sun")
DELIVER_CASES = ("ordinary", "motionblur", "overexposure", "underexposure", "eventlowres", "lidarjitter")


def deliver_keys(filename):
    """(weather, case) of a DELIVER sample from its path, e.g. .../img/cloud/test/MAP_10_point102/045050_rgb_front.png
    with the case given by the split list name (datasets/DELIVER.py:261-300); unknown parts map to None."""
    parts = str(filename).replace("\\", "/").split("/")
    weather = next((p for p in parts if p in DELIVER_WEATHERS), None)
    case = next((c for c in DELIVER_CASES if any(c in p for p in parts)), None)
    return weather, case


class BucketedConfusion:
    """Per-condition evaluation (datasets/DELIVER.py:261-615, apis/test_bs.py:91-163) kept on the device: one [C, C]
    int64 confusion matrix per bucket key plus the global one. add() runs the confusion kernel once per image straight
    into the image's bucket(s); gather() is ONE all-gather of the [K+1, C, C] stack; metrics() applies
    total_area_to_metrics per bucket. Counts are integers, so sharding the images over ranks changes nothing."""

    def __init__(self, num_classes, keys, device, ignore_index=255):
        self.ncls, self.ignore_index = num_classes, ignore_index
        self.keys = ["global"] + [k for k in keys]
        self.index = {k: i for i, k in enumerate(self.keys)}
        self.conf = torch.zeros((len(self.keys), num_classes, num_classes), dtype=torch.int64, device=device)

    def add(self, labels, gt, image_keys=None):
        """labels / gt uint8 [B, H, W] (CUDA); image_keys: per image an iterable of bucket keys (None / unknown keys are
        skipped); every image also counts into 'global'."""
        from . import kernels as K
        B = labels.shape[0]
        for i in range(B):
            ks = ["global"] + [k for k in (image_keys[i] if image_keys is not None else ()) if k in self.index]
            for k in ks:
                K.confusion(labels[i], gt[i], self.ncls, self.ignore_index, out=self.conf[self.index[k]])

    def gather(self, group=None):
        """-> [K+1, C, C] int64 summed over ranks (identical on every rank)."""
        return gather_confusion(self.conf, group)

    def metrics(self, conf_all=None):
        conf_all = self.conf if conf_all is None else conf_all
        return {k: metrics_from_confusion(conf_all[i]) for i, k in enumerate(self.keys)}
