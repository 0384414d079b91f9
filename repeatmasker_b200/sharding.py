"""Multi-GPU partition of the path: units = SimpleBatcher fragments, which are independent, so
fragment b goes to rank b mod G, the library is replicated on every GPU and the per-rank hit lists are gathered on
the host in fragment order.  No data-path collective: torch.distributed is used only to gather the (small) result
records, the way the reference concatenates its per-batch .cat files in order (RepeatMasker:994-1017)."""
import torch.distributed as dist


def my_fragments(n_fragments, rank, world):
    """Indices of the fragments rank `rank` searches (round robin keeps long and short fragments mixed)."""
    return list(range(rank, n_fragments, world))


def gather_hits(local_hits, local_fragment_ids, n_fragments, group=None):
    """local_hits: list (one entry per local fragment, same order as local_fragment_ids) of hit lists.
    Returns, on every rank, the list of hit lists for fragments 0..n_fragments-1."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = [None] * n_fragments
        for f, h in zip(local_fragment_ids, local_hits):
            out[f] = h
        return out
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, (list(local_fragment_ids), list(local_hits)), group=group)
    out = [None] * n_fragments
    for ids, hits in gathered:
        for f, h in zip(ids, hits):
            if out[f] is not None:
                raise RuntimeError(f"fragment {f} was searched by two ranks")
            out[f] = h
    missing = [f for f, h in enumerate(out) if h is None]
    if missing:
        raise RuntimeError(f"fragments {missing[:5]} were searched by no rank")
    return out


def split_hits_by_fragment(hits, names):
    """Group a flat, query-ordered hit list by query record (= fragment) name order."""
    pos = {n: i for i, n in enumerate(names)}
    out = [[] for _ in names]
    for h in hits:
        out[pos[h["qname"]]].append(h)
    return out
