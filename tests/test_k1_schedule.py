"""The schedule of the score pass's column segments (kernels.cuh: seg_claim_decode / seg_block_range), checked on
the CPU through the host-side hook that calls the kernel's own functions.  What keeps the kernel free of deadlock is that
a sub-ticket only ever waits for sub-tickets with a SMALLER claim number: its own previous segment, and -- because the
consumer of block b needs its producer's block b + 1 -- the segments of the stripe above that hold the blocks it reads."""
import pytest

from repeatmasker_b200 import engine


def build_order(chains, n_q):
    """Ticket order as engine.cu builds it for one group of chains: level-major, query fastest."""
    order = [(c, lvl) for lvl in range(max(chains)) for c, ln in enumerate(chains) if lvl < ln]
    pos_of = {cl: i for i, cl in enumerate(order)}
    return order, pos_of


@pytest.mark.parametrize("chains,n_q,n_seg,window,blocks", [
    ([16, 8, 4, 4, 4, 2, 1, 1], 3, 4, 20, [400, 401, 97]),       # partial last window, a record at the cutting limit
    ([5, 4, 4, 3, 1], 4, 3, 1000, [64, 30, 200, 31]),            # one window; records too short to be cut (30, 31 < 3 * (4 + 8))
    ([12] * 3 + [1] * 40, 2, 8, 17, [1000, 999]),                # deep chains, windows smaller than a level
    ([4] * 10, 5, 2, 7, [50, 51, 52, 53, 20]),
])
def test_every_dependency_of_a_sub_ticket_is_claimed_before_it(chains, n_q, n_seg, window, blocks):
    order, pos_of = build_order(chains, n_q)
    n_tickets = len(order) * n_q
    max_level = max(chains) - 1
    seg_min = n_seg * (max_level + 8)
    claim_of, rng = {}, {}
    for claim in range(n_tickets * n_seg):
        t0, s0 = claim, 0
        pos, q = None, None
        # the level belongs to the ticket the claim decodes to: decode first with a dummy record, then with the real one
        _, t, s, _, _ = engine.k1_schedule(n_tickets, n_seg, window, claim, 1 << 20, seg_min, 0)
        pos, q = divmod(t, n_q)
        lvl = order[pos][1]
        work, t2, s2, bs, be = engine.k1_schedule(n_tickets, n_seg, window, claim, blocks[q], seg_min, lvl)
        assert (t2, s2) == (t, s)
        assert (t, s) not in claim_of, "a sub-ticket claimed twice"
        claim_of[(t, s)] = claim
        rng[(t, s)] = (bs, be) if work else None
    assert len(claim_of) == n_tickets * n_seg  # a bijection between claims and (ticket, segment)
    for t in range(n_tickets):
        pos, q = divmod(t, n_q)
        c, lvl = order[pos]
        nb = blocks[q]
        segs = [(s, rng[(t, s)]) for s in range(n_seg) if rng[(t, s)] is not None]
        # the segments with work tile [0, nb) in order, and only the first of a short record has any
        assert segs[0][1][0] == 0 and segs[-1][1][1] == nb
        for (sa, (_, ea)), (sb, (bb, _)) in zip(segs, segs[1:]):
            assert ea == bb and sa < sb
        assert all(be - bs >= 8 for _, (bs, be) in segs) or len(segs) == 1
        if nb < seg_min:
            assert [s for s, _ in segs] == [0]
        for s, (bs, be) in segs:
            if s > 0:  # own previous segment
                assert claim_of[(t, s - 1)] < claim_of[(t, s)]
            if lvl > 0:  # the stripe above: block b of the consumer reads what the producer wrote during its block b + 1
                tp = pos_of[(c, lvl - 1)] * n_q + q
                q_blocks = nb - 1  # the last block only drains the wavefront and reads nothing
                for b in {bs, min(be, q_blocks) - 1}:
                    if b < bs:
                        continue
                    need = b + 1
                    owner = [sp for sp in range(n_seg) if rng[(tp, sp)] and rng[(tp, sp)][0] <= need < rng[(tp, sp)][1]]
                    assert len(owner) == 1
                    assert owner[0] <= s, "a consumer segment needs a LATER segment of its producer"
                    assert claim_of[(tp, owner[0])] < claim_of[(t, s)]


def test_whole_tickets_are_the_identity():
    for claim in (0, 5, 99):
        assert engine.k1_schedule(100, 1, 100, claim, 77, 0, 3) == (1, claim, 0, 0, 77)
    with pytest.raises(ValueError):
        engine.k1_schedule(100, 4, 50, 400, 77, 0, 0)
