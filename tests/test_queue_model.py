"""Index arithmetic of the detection kernel's static ownership schemes, restated in Python (collide_kernels.cu:
traverse_queue, aux_loop, narrow_rest). The CUDA code cannot run here; what can be checked without a GPU is that the
formulas hand every queue slot and every candidate to exactly one owner for any sizes -- a gap or an overlap there is a
lost or a duplicated colliding pair."""
import random

import pytest

COL_WARPS = 24          # kColWarps: 768 threads
AUX_WARPS = 3           # kAuxWarps: 1 control + 2 narrow
TRAV_WARPS = COL_WARPS - AUX_WARPS
NARROW_WARPS = AUX_WARPS - 1
GROUP = 32              # candidates per group


def test_model_constants_match_the_kernel_source():
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oibvh_b200", "csrc",
                            "collide_kernels.cu")).read()
    assert int(re.search(r"#define OIBVH_COL_THREADS (\d+)", src).group(1)) == COL_WARPS * 32
    assert int(re.search(r"#define OIBVH_COL_AUX_WARPS (\d+)", src).group(1)) == AUX_WARPS
    assert "narrow_range(sh, (uint32_t)(g * 32u), 32u, lane)" in src  # GROUP


def slot_owner(slot, ctas):
    """queue slot -> (cta, traversal warp): warp gw owns slots gw, gw + W, gw + 2 W, ... with gw = cta * kTravWarps + warp"""
    gw = slot % (ctas * TRAV_WARPS)
    return divmod(gw, TRAV_WARPS)


@pytest.mark.parametrize("ctas", [1, 2, 148])
def test_every_queue_slot_has_one_owner_and_owners_get_equal_shares(ctas):
    W = ctas * TRAV_WARPS
    n = 5 * W + 17
    seen = {}
    for cta in range(ctas):
        for warp in range(TRAV_WARPS):
            gw, k = cta * TRAV_WARPS + warp, 0
            while gw + k * W < n:  # the slots this warp polls, in this order
                assert (gw + k * W) not in seen
                seen[gw + k * W] = (cta, warp)
                k += 1
    assert sorted(seen) == list(range(n))
    assert all(seen[s] == slot_owner(s, ctas) for s in range(n))
    per_warp = [sum(1 for s in seen.values() if s == (c, w)) for c in range(ctas) for w in range(TRAV_WARPS)]
    assert max(per_warp) - min(per_warp) <= 1


def narrow_phase_model(ctas, n_cand, progress):
    """progress[(cta, a)] = groups narrow warp a of that CTA tested while the traversal ran (only whole groups below the
    tail it saw). Returns how often every candidate is tested: by its narrow warp first, then by the CTA's warps that
    share out the backlog (narrow_rest)."""
    NW = ctas * NARROW_WARPS
    hits = [0] * n_cand
    step = COL_WARPS // NARROW_WARPS
    for cta in range(ctas):
        backlog = []
        for a in range(NARROW_WARPS):
            nw = cta * NARROW_WARPS + a
            k = 0
            while k < progress.get((cta, a), 0) and ((nw + k * NW) + 1) * GROUP <= n_cand:
                for i in range((nw + k * NW) * GROUP, (nw + k * NW + 1) * GROUP):
                    hits[i] += 1
                k += 1
            backlog.append(k)
        for warp in range(COL_WARPS):
            a = warp % NARROW_WARPS
            nw = cta * NARROW_WARPS + a
            k = backlog[a] + warp // NARROW_WARPS
            while (nw + k * NW) * GROUP < n_cand:
                first = (nw + k * NW) * GROUP
                for i in range(first, min(first + GROUP, n_cand)):
                    hits[i] += 1
                k += step
    return hits


@pytest.mark.parametrize("ctas", [1, 3, 148])
def test_every_candidate_is_tested_exactly_once(ctas):
    rng = random.Random(ctas)
    assert COL_WARPS % NARROW_WARPS == 0
    for n_cand in (0, 1, 31, 32, 33, 1000, 19090, 70001):
        for trial in range(3):
            progress = {(c, a): rng.choice([0, 0, 1, 2, 5, 10 ** 6]) for c in range(ctas) for a in range(NARROW_WARPS)}
            hits = narrow_phase_model(ctas, n_cand, progress)
            assert all(h == 1 for h in hits), (ctas, n_cand, trial)


def test_termination_count_includes_one_seeding_token_per_cta():
    """the stop flag is raised when finished == pushed + ctas: every CTA's seeding is an item in flight that is never
    pushed (queue_retire). With any interleaving of pushes and retirements the condition first holds at the very end."""
    rng = random.Random(7)
    ctas = 5
    for trial in range(50):
        pushed = finished = 0
        tokens = ctas       # not yet retired
        in_flight = []      # items taken but not finished; each will push some children first
        queue = 0           # pushed, not yet taken
        budget = 200
        while tokens or in_flight or queue:
            moves = []
            if tokens:
                moves.append("seed")
            if queue:
                moves.append("take")
            if in_flight:
                moves.append("finish")
            m = rng.choice(moves)
            if m == "seed":       # a CTA pushes its seeds, THEN retires its token
                n = rng.randint(0, 3)
                pushed += n
                queue += n
                tokens -= 1
                finished += 1
            elif m == "take":
                queue -= 1
                in_flight.append(rng.randint(0, 2) if budget > 0 else 0)
            else:                 # children are pushed BEFORE the item is retired
                n = in_flight.pop(rng.randrange(len(in_flight)))
                budget -= n
                pushed += n
                queue += n
                finished += 1
            done = finished == pushed + ctas
            assert done == (not tokens and not in_flight and not queue)
        assert finished == pushed + ctas
