"""Picture-level sharding of the hot path over GPUs (one process per GPU).

Motion estimation, the transform pass and deblocking of one picture depend only on that picture and its reference
pictures; a closed mini-GOP (the reference's prediction structure, EbPredictionStructure.c: 16 pictures at 4
hierarchical levels, every picture referencing only its own mini-GOP's anchors and the previous base-layer picture)
is therefore the unit that moves between GPUs: rank r owns mini-GOPs r, r+world, r+2*world, ... together with the one
base-layer picture that precedes each of them.  No collective runs on the data path; torch.distributed is used only to
agree on the step time (max over ranks) and to total the pictures processed."""
import torch
import torch.distributed as dist

from . import presets


def minigop_encode_order(levels=4):
    """Display offsets 1..2^levels of one mini-GOP in encode (= decode) order: base-layer picture first, then each
    layer's pictures after both of their references (EbPictureDecisionProcess.c orders by the prediction structure)."""
    n = 1 << levels
    order = [n]

    def rec(lo, hi):
        if hi - lo < 2:
            return
        mid = (lo + hi) // 2
        order.append(mid)
        rec(lo, mid)
        rec(mid, hi)
    rec(0, n)
    return order


def minigop_refs(offset, levels=4):
    """(list0, list1) display offsets (relative to the mini-GOP start; 0 = previous base-layer picture) of the two
    reference pictures of the picture at 'offset' (1..2^levels); the base-layer picture references the previous one twice."""
    n = 1 << levels
    if offset == n:
        return 0, 0
    step = offset & -offset
    return offset - step, offset + step


def temporal_layer(offset, levels=4):
    n = 1 << levels
    assert 1 <= offset <= n
    tl = presets.MINIGOP16_TL[offset - 1] if levels == 4 else levels - ((offset & -offset).bit_length() - 1)
    return tl


def shard_minigops(n_pictures, world, rank, levels=4):
    """Pictures (display numbers, picture 0 = the intra picture) owned by 'rank': whole mini-GOPs dealt round-robin.
    Returns a list of (first_display_number, count) runs; every picture 1..n_pictures-1 belongs to exactly one rank."""
    n = 1 << levels
    runs = []
    g = rank
    while 1 + g * n < n_pictures:
        first = 1 + g * n
        runs.append((first, min(n, n_pictures - first)))
        g += world
    return runs


def pictures_needed(runs):
    """Display numbers that must be resident on the owning GPU: the run's pictures plus the preceding base-layer picture."""
    need = set()
    for first, count in runs:
        need.add(first - 1)
        need.update(range(first, first + count))
    return sorted(need)


def agree_step(ms_local, units_local, device=None):
    """(max over ranks of the step time, total units over ranks).  Works on any initialised backend (nccl on GPUs, gloo on
    CPU); a single process without a process group is its own world."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(ms_local), int(units_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([int(units_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), int(u[0])
