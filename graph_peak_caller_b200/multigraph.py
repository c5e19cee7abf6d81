"""The one cross-chromosome step of the path: the p -> q mapping is built from
the p-values of ALL chromosome graphs (reference
PToQValuesMapper.from_files, sparsepvalues.py:76-93, joined through
``*_pvalues_*.npy`` files on disk; multiplegraphscallpeaks.py:122-124).

With one process per GPU and chromosomes sharded over ranks, every rank
reduces its chromosomes to a table of (distinct p, base pairs) and the tables
are exchanged with one all-gather over NCCL (NVLink / NVSwitch); every rank
then builds the identical global q table.  Counts are integers, so the result
does not depend on the order of the exchange."""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def gather_p_tables(table_p, table_count, total_bp, group=None):
    """table_p float64[D], table_count int64[D] (torch tensors, CUDA for NCCL,
    CPU for gloo), total_bp: local sum of track sizes.  Returns the
    concatenation over all ranks and the global total."""
    import torch
    d = _dist()
    if d is None or d.get_world_size(group) == 1:
        return table_p, table_count, int(total_bp)
    ws = d.get_world_size(group)
    meta = torch.tensor([table_p.numel(), int(total_bp)], dtype=torch.int64,
                        device=table_p.device)
    metas = [torch.empty_like(meta) for _ in range(ws)]
    d.all_gather(metas, meta, group=group)
    metas = torch.stack(metas).cpu().numpy()
    sizes, totals = metas[:, 0], metas[:, 1]
    width = int(sizes.max())
    buf = torch.zeros((2, max(width, 1)), dtype=torch.int64, device=table_p.device)
    n = table_p.numel()
    buf[0, :n] = table_p.view(torch.int64)        # bit pattern of the doubles
    buf[1, :n] = table_count
    bufs = [torch.empty_like(buf) for _ in range(ws)]
    d.all_gather(bufs, buf, group=group)
    ps = torch.cat([b[0, :int(s)] for b, s in zip(bufs, sizes)]).view(torch.float64)
    cs = torch.cat([b[1, :int(s)] for b, s in zip(bufs, sizes)])
    return ps, cs, int(totals.sum())


def assign_chromosomes(sizes, n_ranks):
    """Longest-processing-time bin packing of chromosome graphs onto ranks
    (SURVEY.md §8e).  ``sizes``: cost per chromosome.  Returns a list of index
    lists, one per rank."""
    order = np.argsort(-np.asarray(sizes, dtype=np.float64), kind="stable")
    load = np.zeros(n_ranks)
    bins = [[] for _ in range(n_ranks)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += sizes[i]
    return bins
