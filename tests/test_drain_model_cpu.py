"""A model of the drain of the persistent traversal kernel (csrc/trace.cu, "shared walks once the pool is empty"): the
argument that the record does not depend on how a ray's stack is split between lanes, checked on random trees.

The model keeps what the argument uses and nothing else: a binary tree whose inner nodes carry a lower bound of the hit
distances below them (the entry distance of the slab test), leaves that hit at some t or miss, the closest hit as the
maximum of the total order (smaller t, then larger leaf index), pruning of entries whose bound exceeds a limit derived
from the best t known to the lane (with the kernel's slack: never below the t itself), an owner lane plus helpers that
take every other entry of a walking lane's stack, publish their best hit to a word the whole group prunes with, and the
owner taking that word once every helper is back.  Any interleaving must return what one lane alone returns."""
import random

import pytest

NONE = -1


def make_tree(rng, n_leaves):
    """random binary tree in DFS pre-order: nodes[i] = (lower_bound, left, right) or (t_hit or None, leaf_id)"""
    nodes = []

    def build(lo, hi, bound):
        idx = len(nodes)
        nodes.append(None)
        if hi - lo == 1:
            t = bound + rng.choice([0.0, 0.0, rng.random()]) if rng.random() < 0.5 else None   # ties in t are common
            nodes[idx] = ("leaf", t, idx)
            return idx
        mid = rng.randint(lo + 1, hi - 1)
        bl = bound + rng.choice([0.0, rng.random() * 0.5])
        br = bound + rng.choice([0.0, rng.random() * 0.5])
        l = build(lo, mid, bl)
        r = build(mid, hi, br)
        nodes[idx] = ("inner", (bl, l), (br, r))
        return idx

    build(0, n_leaves, 0.0)
    return nodes


def better(t, node, best):
    bt, bn = best
    return bn == NONE or t < bt or (t == bt and node > bn)


def limit_of(t, slack):
    return float("inf") if t is None else t * (1.0 + slack) + slack     # >= t: a tie is never pruned


class Lane:
    def __init__(self, stack, best, limit):
        self.stack, self.best, self.limit = list(stack), best, limit

    def step(self, nodes, slack):
        """pop one entry, expand it; returns False when the stack is dry"""
        while self.stack:
            bound, idx = self.stack.pop()
            if bound > self.limit:
                continue
            node = nodes[idx]
            if node[0] == "leaf":
                _, t, leaf = node
                if t is not None and better(t, leaf, self.best):
                    self.best = (t, leaf)
                    self.limit = min(self.limit, limit_of(t, slack))
            else:
                (bl, l), (br, r) = node[1], node[2]
                near, far = ((bl, l), (br, r)) if bl <= br else ((br, r), (bl, l))
                self.stack.append(far)
                self.stack.append(near)
            return True
        return False


def walk_alone(nodes, slack):
    lane = Lane([(0.0, 0)], (float("inf"), NONE), float("inf"))
    while lane.step(nodes, slack):
        pass
    return lane.best


def walk_shared(nodes, slack, rng, lanes=8):
    owner = Lane([(0.0, 0)], (float("inf"), NONE), float("inf"))
    helpers, published = [], (float("inf"), NONE)
    while True:
        group = [owner] + helpers
        # idle lanes take every other entry of a walking lane's stack (bottom first), dead entries dropped
        if len(group) < lanes and rng.random() < 0.5:
            donors = [g for g in group if len(g.stack) >= 2]
            if donors:
                d = rng.choice(donors)
                taken = [e for e in d.stack[0::2] if not e[0] > d.limit]
                d.stack = [e for e in d.stack[1::2] if not e[0] > d.limit]
                helpers.append(Lane(taken, d.best, d.limit))
        # some lanes move (any subset, any order)
        for g in rng.sample(group + helpers[len(group) - 1:], k=len(group)):
            if rng.random() < 0.7:
                g.step(nodes, slack)
        # the lanes of the ray publish, prune with the group's best, dry helpers fall idle
        for g in [owner] + helpers:
            if g.best[1] != NONE and better(g.best[0], g.best[1], published):
                published = g.best
        if published[1] != NONE:
            for g in [owner] + helpers:
                g.limit = min(g.limit, limit_of(published[0], slack))
        helpers = [h for h in helpers if h.stack]
        if not owner.stack and not helpers:
            if published[1] != NONE and better(published[0], published[1], owner.best):
                owner.best = published
            return owner.best


@pytest.mark.parametrize("seed", range(40))
def test_any_split_of_the_stack_returns_the_one_lane_record(seed):
    rng = random.Random(seed)
    nodes = make_tree(rng, rng.randint(1, 400))
    slack = rng.choice([0.0, 1e-3, 0.2])
    alone = walk_alone(nodes, slack)
    # the definition: minimum t over all hitting leaves, the largest leaf index among equal t
    hits = [(n[1], n[2]) for n in nodes if n[0] == "leaf" and n[1] is not None]
    expect = (float("inf"), NONE)
    for t, leaf in hits:
        if better(t, leaf, expect):
            expect = (t, leaf)
    assert alone == expect
    for trial in range(5):
        assert walk_shared(nodes, slack, random.Random(seed * 100 + trial)) == expect
