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


_capacity = 1 << 16      # entries per rank in the exchange buffer (grown on demand)


def gather_p_tables(table_p, table_count, total_bp, group=None):
    """table_p float64[D], table_count int64[D] (torch tensors, CUDA for NCCL,
    CPU for gloo), total_bp: local sum of track sizes.  Returns the
    concatenation over all ranks and the global total.

    One collective: every rank contributes a fixed-capacity buffer whose first
    column is a header (entries, base pairs); the capacity only grows -- on all
    ranks alike, they read the same headers -- when a table does not fit."""
    global _capacity
    import torch
    d = _dist()
    if d is None or d.get_world_size(group) == 1:
        return table_p, table_count, int(total_bp)
    ws = d.get_world_size(group)
    n = table_p.numel()
    while True:
        cap = _capacity
        buf = torch.empty((2, cap + 1), dtype=torch.int64, device=table_p.device)
        buf[:, 0] = torch.tensor([n, int(total_bp)], dtype=torch.int64)
        m = min(n, cap)
        buf[0, 1:m + 1] = table_p.view(torch.int64)[:m]        # bit pattern of the doubles
        buf[1, 1:m + 1] = table_count[:m]
        out = torch.empty((ws, 2, cap + 1), dtype=torch.int64, device=table_p.device)
        if d.get_backend(group) == "nccl":
            d.all_gather_into_tensor(out, buf, group=group)
        else:                                  # gloo (CPU tests): list form
            d.all_gather([out[r] for r in range(ws)], buf, group=group)
        head = out[:, :, 0].cpu().numpy()
        sizes, totals = head[:, 0], head[:, 1]
        if int(sizes.max()) <= cap:
            break
        while _capacity < int(sizes.max()):
            _capacity *= 2
    ps = torch.cat([out[r, 0, 1:int(k) + 1] for r, k in enumerate(sizes)]).view(torch.float64)
    cs = torch.cat([out[r, 1, 1:int(k) + 1] for r, k in enumerate(sizes)])
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
