"""Point-set preprocessing: centring and scaling into the torus cell [-1/4, 1/4]^d.

Behaviour of reference `torch_nfft/utils.py:6-99`.  The batched variants use
`Tensor.scatter_reduce_` (amin / amax) instead of the optional `torch_scatter` dependency the
reference needs (utils.py:19-28, 67-75).
"""
import torch


def _both(batch, source_batch, target_batch):
    return (batch, batch) if batch is not None else (source_batch, target_batch)


def _segment_extreme(values, index, num_segments, mode):
    """Per-segment min or max over dim 0 of `values` ([n] or [n, d])."""
    shape = (num_segments,) + tuple(values.shape[1:])
    init = float("inf") if mode == "amin" else float("-inf")
    out = torch.full(shape, init, dtype=values.dtype, device=values.device)
    idx = index if values.dim() == 1 else index[:, None].expand_as(values)
    return out.scatter_reduce_(0, idx, values, reduce=mode, include_self=True)


def _num_segments(*batches):
    return max(int(b.max().item()) + 1 for b in batches if b is not None)


def compute_points_center(sources, targets=None, source_batch=None, target_batch=None, /, batch=None):
    """Centre of the bounding box of each point set (sources and targets together)."""
    source_batch, target_batch = _both(batch, source_batch, target_batch)
    sets = [(sources, source_batch)] + ([(targets, target_batch)] if targets is not None else [])
    if source_batch is None:
        lo = torch.stack([p.min(dim=0).values for p, _ in sets]).min(dim=0).values
        hi = torch.stack([p.max(dim=0).values for p, _ in sets]).max(dim=0).values
    else:
        nseg = _num_segments(*(b for _, b in sets))
        lo = torch.stack([_segment_extreme(p, b, nseg, "amin") for p, b in sets]).min(dim=0).values
        hi = torch.stack([_segment_extreme(p, b, nseg, "amax") for p, b in sets]).max(dim=0).values
    return 0.5 * (lo + hi)


def shift_points_by_center(sources, targets=None, source_batch=None, target_batch=None, /, batch=None):
    """Translate every point set so that its bounding-box centre is the origin."""
    source_batch, target_batch = _both(batch, source_batch, target_batch)
    center = compute_points_center(sources, targets, source_batch, target_batch)
    pick = lambda b: center if b is None else center[b]
    return sources - pick(source_batch), (None if targets is None else targets - pick(target_batch))


def _point_norms(points, norm):
    if norm == "euclidean":
        return points.pow(2).sum(dim=1).sqrt()
    if norm == "infinity":
        return points.abs().max(dim=1).values
    raise ValueError(f"scale_points_by_norm received unknown norm: {norm}")


def compute_points_radius(sources, targets=None, source_batch=None, target_batch=None, /, batch=None,
                          norm="euclidean"):
    """Largest point norm per point set: a float without batches, a [batch_size] tensor with."""
    source_batch, target_batch = _both(batch, source_batch, target_batch)
    sets = [(sources, source_batch)] + ([(targets, target_batch)] if targets is not None else [])
    if source_batch is None:
        return max(_point_norms(p, norm).max().item() for p, _ in sets)
    nseg = _num_segments(*(b for _, b in sets))
    per_set = [_segment_extreme(_point_norms(p, norm), b, nseg, "amax") for p, b in sets]
    return torch.stack(per_set).max(dim=0).values


def scale_points_by_norm(sources, targets=None, source_batch=None, target_batch=None, /, batch=None, factor=1,
                         norm="euclidean"):
    """Scale every point set so that its largest point norm becomes `factor`."""
    source_batch, target_batch = _both(batch, source_batch, target_batch)
    scale = factor / compute_points_radius(sources, targets, source_batch, target_batch, norm=norm)
    pick = lambda b: scale if b is None else scale[b, None]
    return sources * pick(source_batch), (None if targets is None else targets * pick(target_batch))
