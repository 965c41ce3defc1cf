"""Static schedule of the patch-grid sampler over ALL stages of the cascade (reference dependency rule
sample_ultra_res.py:99-107 + the stage order of :264-270).

Unit of work: (stage u, patch k).  It depends on the SAME stage of the up-to-three neighbours (i-1,j), (i,j+o), (i-1,j+o) that
exist in the grid, and on stage u-1 of the same patch -- nothing else.  The reference runs the three stages one after the
other over the whole grid (stage-major); here one dependency-driven list schedule covers every stage, so the 64^2 / 256^2 work
of later anti-diagonals runs while earlier anti-diagonals are already in their 1024^2 stage ("pipelined wavefront"): the
launch-bound base-stage chain (41 anti-diagonals x 1024 steps on a 21 x 21 grid) hides under the tensor-bound 1024^2 stage
instead of preceding it.

The schedule is computed identically on every rank from (grid, orientation, world, cost table): an event-driven simulation
assigns batches of ready units to ranks; each rank then executes its own batches in simulated-start order.  Because every
dependency finishes (in simulated time) before its consumer starts, the per-rank orders are consistent with one global
order: ranks that wait on each other's border strips cannot deadlock.  Results never depend on the schedule (counter-based
per-patch noise, batch-invariant kernels), so any cost table gives the same image.
"""
from __future__ import annotations

import json
import math
import os

# ms per sampling step of a batch of B patches, per stage (B200, round-2 measurements of profiles/sweep.py; override with
# kidney_diffusion_b200/cost_model.json or the KD_COST_MODEL environment variable).  Only relative values matter.
DEFAULT_STEP_MS = {
    1: {1: 6.3, 2: 6.4, 4: 6.7, 8: 7.4, 16: 9.1, 32: 12.8, 64: 20.8},
    2: {1: 4.5, 2: 4.6, 4: 5.4, 8: 6.8, 16: 10.2, 32: 18.4, 64: 35.4},
    3: {1: 15.0, 2: 26.9, 4: 50.8, 8: 97.9, 16: 192.6, 32: 383.5},
}
# batch sizes a plan may use (bounds the number of captured CUDA graphs): fine for the small launch-bound stages, where taking a
# whole anti-diagonal (up to 21 patches on the 21 x 21 grid) in ONE batch matters; coarse for the 1024^2 stage (3.3 GB per patch)
ALLOWED_BATCH = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32)
ALLOWED_BATCH_SMALL = (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 18, 21, 24, 28, 32)
DEFAULT_MAX_BATCH = {1: 32, 2: 16, 3: 16}                # per stage (64^2 / 256^2 / 1024^2): memory- and latency-driven caps
FULL_STEPS = {1: 1024, 2: 256, 3: 256}                   # train_ultra_res_v_param.py:86


def load_cost_table():
    path = os.environ.get("KD_COST_MODEL") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "cost_model.json")
    if os.path.exists(path):
        with open(path) as fh:
            raw = json.load(fh)
        return {int(u): {int(b): float(v) for b, v in t.items()} for u, t in raw.items() if str(u).isdigit()}
    return DEFAULT_STEP_MS


def step_ms(table, stage, B):
    """Piecewise-linear in B between measured points, linear extrapolation beyond the last two."""
    t = table[stage]
    bs = sorted(t)
    if B in t:
        return t[B]
    if B <= bs[0]:
        return t[bs[0]]
    for lo, hi in zip(bs, bs[1:]):
        if lo < B < hi:
            return t[lo] + (t[hi] - t[lo]) * (B - lo) / (hi - lo)
    lo, hi = bs[-2], bs[-1]
    return t[hi] + (t[hi] - t[lo]) * (B - hi) / (hi - lo)


def neighbours(pos, orientation):
    i, j = pos
    return dict(above=(i - 1, j), side=(i, j + orientation), corner=(i - 1, j + orientation))


def dependents(pos, orientation):
    """Patches that use `pos` as their above / side / corner neighbour, and which strip of `pos` each needs."""
    i, j = pos
    return dict(above=(i + 1, j), side=(i, j - orientation), corner=(i + 1, j - orientation))


class Batch:
    __slots__ = ("stage", "patches", "rank", "start", "finish")

    def __init__(self, stage, patches, rank, start, finish):
        self.stage, self.patches, self.rank, self.start, self.finish = stage, patches, rank, start, finish

    def __repr__(self):
        return f"Batch(u{self.stage} {self.patches} @r{self.rank} {self.start:.1f}-{self.finish:.1f})"


class Plan:
    def __init__(self, batches, owner, makespan, stages, world, policy):
        self.batches, self.owner, self.makespan, self.stages, self.world, self.policy = batches, owner, makespan, stages, world, policy

    def for_rank(self, rank):
        return [b for b in self.batches if b.rank == rank]

    def batch_sizes(self):
        out = {}
        for b in self.batches:
            out.setdefault(b.stage, set()).add(len(b.patches))
        return {u: sorted(v) for u, v in out.items()}

    def busy_fraction(self):
        busy = sum(b.finish - b.start for b in self.batches)
        return busy / (self.world * self.makespan) if self.makespan > 0 else 1.0


def _floor_allowed(n, cap, stage=3):
    best = 1
    for b in (ALLOWED_BATCH if stage >= 3 else ALLOWED_BATCH_SMALL):
        if b <= n and b <= cap:
            best = b
    return best


def _simulate(patch_pos, orientation, world, stages, unit_cost, max_batch, pack_stages, cost_fn):
    """One list-scheduling pass.  unit_cost[u]: cost of a B=1 batch (priorities); cost_fn(u, B): cost of a batch."""
    n = len(patch_pos)
    index = {p: k for k, p in enumerate(patch_pos)}
    stages = list(stages)
    deps, users = {}, {}
    for u in stages:
        for k, p in enumerate(patch_pos):
            d = [(u, index[nb]) for nb in neighbours(p, orientation).values() if nb in index]
            if (u - 1) in stages:
                d.append((u - 1, k))
            deps[(u, k)] = d
            for t in d:
                users.setdefault(t, []).append((u, k))
    # priority = length (in B=1 cost) of the longest chain of dependents hanging off a unit, itself included
    prio = {}
    order = sorted(deps, key=lambda t: (-t[0], -patch_pos[t[1]][0], orientation * patch_pos[t[1]][1]))
    for t in order:  # later stages first; bottom rows first; along the dependency direction first => users are done before t
        prio[t] = unit_cost[t[0]] + max((prio[x] for x in users.get(t, ())), default=0.0)
    missing = {t: len(d) for t, d in deps.items()}
    ready = {t: 0.0 for t, m in missing.items() if m == 0}
    finish, owner, batches = {}, {}, []
    free_at = [0.0] * world
    eps = 1e-9
    left = len(deps)
    while left:
        r = min(range(world), key=lambda q: (free_at[q], q))
        t_now = free_at[r]
        avail = [t for t, rt in ready.items() if rt <= t_now + eps]
        if not avail:
            assert ready, "dependency cycle in the patch grid"
            free_at[r] = min(rt for rt in ready.values())
            continue
        avail.sort(key=lambda t: (-prio[t], t[0], patch_pos[t[1]]))
        u = avail[0][0]
        cand = [t for t in avail if t[0] == u]
        if u in pack_stages:
            cap = max_batch[u]
        else:  # share what is ready among the ranks that are idle right now
            idle = sum(1 for q in range(world) if free_at[q] <= t_now + eps)
            cap = min(max_batch[u], max(1, -(-len(cand) // idle)))
        take = cand[:_floor_allowed(len(cand), cap, u)]
        end = t_now + cost_fn(u, len(take))
        batches.append(Batch(u, [t[1] for t in take], r, t_now, end))
        free_at[r] = end
        for t in take:
            finish[t], owner[t] = end, r
            del ready[t]
            left -= 1
            for x in users.get(t, ()):
                missing[x] -= 1
                if missing[x] == 0:
                    ready[x] = max(finish[d] for d in deps[x])
    makespan = max(free_at) if batches else 0.0
    batches.sort(key=lambda b: (b.start, b.rank))
    return batches, owner, makespan


def build_plan(patch_pos, orientation, world, stages=(1, 2, 3), steps=None, resample=1, max_batch=None, table=None, policy=None):
    """Pipelined wavefront plan for `stages` (subset of 1..3, consecutive) of one magnification level.

    steps: {stage: sampling steps}; resample: inner RePaint iterations per step; max_batch: int (all stages) or {stage: cap}.
    policy: None = try the packing variants below and keep the one with the smallest simulated makespan."""
    patch_pos = [tuple(p) for p in patch_pos]
    stages = tuple(sorted(stages))
    table = table or load_cost_table()
    steps = {**FULL_STEPS, **(steps or {})}
    mb = dict(DEFAULT_MAX_BATCH)
    if isinstance(max_batch, int):
        mb = {u: max_batch for u in (1, 2, 3)}
    elif max_batch:
        mb.update(max_batch)
    n_inner = {u: steps[u] * max(1, resample) for u in stages}
    cost_fn = lambda u, B: n_inner[u] * step_ms(table, u, B) * 1e-3
    unit = {u: cost_fn(u, 1) for u in stages}
    # "pack" a stage = give one rank everything that is ready (the launch-bound 64^2 / 256^2 stages cost almost the same for
    # B = 1 and B = 16, so spreading them over ranks wastes GPU-seconds the 1024^2 stage needs)
    if policy is not None:
        variants = [(tuple(policy), mb.get(3, 16))]
    else:  # small search: which stages to pack x a cap on the 1024^2 batch (smaller batches shorten the wavefront's critical path)
        caps3 = sorted({min(mb.get(3, 16), c) for c in (16, 4, 3, 2)}, reverse=True) if (3 in stages and world > 1) else [mb.get(3, 16)]
        variants = [(pack, c3) for pack in ((), (1,), (1, 2), (1, 2, 3)) for c3 in caps3]
    best = None
    for pack, cap3 in variants:
        pack = tuple(u for u in pack if u in stages)
        batches, owner, makespan = _simulate(patch_pos, orientation, world, stages, unit, {**mb, 3: cap3}, pack, cost_fn)
        if best is None or makespan < best[2] - 1e-9:
            best = (batches, owner, makespan, (pack, cap3))
    return Plan(best[0], best[1], best[2], stages, world, best[3])


def simulate_makespan(plan, patch_pos, orientation, steps, resample, table):
    """Re-time a fixed plan (same batches, same per-rank order) under another cost table / step count: the critical-path
    recurrence finish(b) = max(finish of the previous batch on the rank, finish of every dependency) + cost(b)."""
    patch_pos = [tuple(p) for p in patch_pos]
    index = {p: k for k, p in enumerate(patch_pos)}
    fin, last = {}, {}
    makespan = 0.0
    for b in plan.batches:  # sorted by simulated start: a valid topological order
        t0 = last.get(b.rank, 0.0)
        for k in b.patches:
            for nb in neighbours(patch_pos[k], orientation).values():
                if nb in index:
                    t0 = max(t0, fin[(b.stage, index[nb])])
            if (b.stage - 1) in plan.stages:
                t0 = max(t0, fin[(b.stage - 1, k)])
        end = t0 + steps[b.stage] * max(1, resample) * step_ms(table, b.stage, len(b.patches)) * 1e-3
        for k in b.patches:
            fin[(b.stage, k)] = end
        last[b.rank] = end
        makespan = max(makespan, end)
    return makespan
