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
