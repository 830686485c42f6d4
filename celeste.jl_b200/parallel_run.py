"""Source partitioning and multi-GPU sharding -- the `ParallelRun` side of the hot path.

Mirrors the partition logic of the reference (host-side, integer work):
  src/partition.jl:3-16      union_find!
  src/partition.jl:37-73     compute_connected_components
  src/partition.jl:173-236   partition_cyclades_dynamic  (batches of mutually non-conflicting components)
  src/partition.jl:250-273   partition_equally
  src/ParallelRun.jl:45-56   estimate_time (sum of active pixels) + load_balance_across_threads (greedy)
and extends the reference's thread-level rule to GPU ranks: one process per GPU, connected
components dealt by cost to the least-loaded rank, NO collective on the data path.  The only collectives
are optional and tiny: a sum of the per-rank ELBO (1 double) and an all-gather of updated variational
parameters (44 doubles per source) at a Cyclades batch barrier (ParallelRun.jl:321-324); they go through
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def union_find(i: int, tree: List[int]) -> int:
    """partition.jl:3-16 (with path compression)."""
    root = i
    while tree[i] != i:
        i = tree[i]
    while root != i:
        nxt = tree[root]
        tree[root] = i
        root = nxt
    return i


def compute_connected_components(sources: Sequence[int], neighbor_map: Dict[int, List[int]]) -> List[List[int]]:
    """partition.jl:37-73 for one batch: connected components of the conflict graph restricted to `sources`.
    Components are returned in order of their first member (deterministic, unlike a Julia Dict)."""
    index = {s: k for k, s in enumerate(sources)}
    tree = list(range(len(sources)))
    for k, s in enumerate(sources):
        target = union_find(k, tree)
        for nb in neighbor_map[s]:
            if nb in index:
                comp = union_find(index[nb], tree)
                tree[comp] = target
    comps: Dict[int, List[int]] = {}
    for k, s in enumerate(sources):
        comps.setdefault(union_find(k, tree), []).append(s)
    return list(comps.values())


def partition_cyclades_dynamic(target_sources: Sequence[int], neighbor_map: Dict[int, List[int]],
                               batch_size: int = 60, seed: int = 42) -> List[List[List[int]]]:
    """partition.jl:173-236: shuffle, cut into batches of `batch_size`, split each batch into connected
    components.  Returns [batch][component][source].  Sources in different components of one batch never
    share pixels, so they can be optimised concurrently (by threads there, by GPU ranks here)."""
    rng = np.random.default_rng(seed)          # srand(42), ParallelRun.jl:142
    sources = list(target_sources)
    rng.shuffle(sources)
    out = []
    for start in range(0, len(sources), batch_size):
        out.append(compute_connected_components(sources[start:start + batch_size], neighbor_map))
    assert sum(len(c) for b in out for c in b) == len(target_sources)
    return out


def partition_equally(n_parts: int, n_sources: int) -> List[List[int]]:
    """partition.jl:250-273 (0-based source indices; one batch)."""
    per = n_sources // n_parts
    return [list(range(p * per, n_sources if p == n_parts - 1 else (p + 1) * per)) for p in range(n_parts)]


def estimate_time(patches_row) -> int:
    """ParallelRun.jl:45-47: the cost model is the number of active pixels over all images."""
    return int(sum(int(p.active_pixel_bitmap.sum()) for p in patches_row))


def load_balance(costs: Sequence[float], n_ranks: int) -> List[List[int]]:
    """ParallelRun.jl:49-56 generalised: items in the given order, each to the currently least-loaded rank.
    Returns the item indices per rank."""
    load = np.zeros(n_ranks)
    out: List[List[int]] = [[] for _ in range(n_ranks)]
    for i, c in enumerate(costs):
        r = int(np.argmin(load))
        load[r] += c
        out[r].append(i)
    return out


def shard_batch(components: List[List[int]], cost_of: Callable[[int], float], n_ranks: int) -> List[List[int]]:
    """One Cyclades batch -> sources per rank; a component never straddles ranks (no conflicts across GPUs).
    Components are taken heaviest first."""
    ccost = [sum(cost_of(s) for s in comp) for comp in components]
    order = sorted(range(len(components)), key=lambda i: -ccost[i])
    assign = load_balance([ccost[i] for i in order], n_ranks)
    return [[s for i in idx for s in components[order[i]]] for idx in assign]


def shard_sources(costs: Sequence[float], rank: int, world: int) -> List[int]:
    """Throughput configs (every source evaluated once with frozen neighbours): plain cost-balanced split,
    heaviest first.  Returns this rank's sorted source indices."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    assign = load_balance([costs[i] for i in order], world)
    return sorted(int(order[i]) for i in assign[rank])


# ------------------------------------------------------------------ collectives (optional, tiny)
def allreduce_elbo(local_sum: float, group=None) -> float:
    """Global ELBO = sum over ranks of the local ELBO sums (1 double)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(local_sum)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([local_sum], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def allgather_vp(local_ids: Sequence[int], local_vp: np.ndarray, n_sources: int, group=None) -> np.ndarray:
    """Exchange updated variational parameters at a batch barrier (ParallelRun.jl:321-324): every rank
    contributes the 44-vectors of the sources it owns and receives the full 44 x n_sources table
    (columns it does not learn about stay NaN)."""
    import torch
    import torch.distributed as dist
    table = np.full((44, n_sources), np.nan)
    table[:, list(local_ids)] = local_vp
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return table
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    world = dist.get_world_size(group)
    ids = torch.full((n_sources,), -1, dtype=torch.int64, device=dev)
    ids[:len(local_ids)] = torch.as_tensor(list(local_ids), dtype=torch.int64, device=dev)
    vals = torch.zeros((n_sources, 44), dtype=torch.float64, device=dev)
    vals[:len(local_ids)] = torch.as_tensor(np.ascontiguousarray(local_vp.T), device=dev)
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_vals = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(all_ids, ids, group=group)
    dist.all_gather(all_vals, vals, group=group)
    for i, v in zip(all_ids, all_vals):
        i = i.cpu().numpy()
        v = v.cpu().numpy()
        k = i >= 0
        table[:, i[k]] = v[k].T
    return table


def evaluate_sharded(n_sources: int, costs: Sequence[float], evaluate: Callable[[List[int]], Dict[str, np.ndarray]],
                     rank: int, world: int, reduce_elbo: bool = True):
    """Evaluate every source once (configs[2]/[3]): this rank's shard through `evaluate(source_indices)`
    (the CUDA plan in production; any callable with the same contract in tests).  Returns
    (my_indices, outputs, global_elbo or None)."""
    mine = shard_sources(costs, rank, world)
    out = evaluate(mine)
    total = allreduce_elbo(float(np.sum(out["v"]))) if reduce_elbo else None
    return mine, out, total


# ------------------------------------------------------------------ inference drivers (ParallelRun.jl:135-300, 468-608)
class OptimizedSource:
    """ParallelRun.jl:425-430."""

    def __init__(self, init_ra, init_dec, vs, is_sky_bad=False):
        self.init_ra, self.init_dec, self.vs, self.is_sky_bad = init_ra, init_dec, vs, is_sky_bad


def _task_rows(target_sources, neighbor_map):
    rows, act = [], []
    for s in target_sources:
        rows.append([s + 1] + [n + 1 for n in neighbor_map[s]])       # target first (ParallelRun.jl:242, 485)
        act.append([1])
    return rows, act


def one_node_single_infer(catalog, patches, target_sources, neighbor_map, images, field=None, include_kl=True,
                          rank=0, world=1, max_iters=50):
    """ParallelRun.one_node_single_infer (:546-608) with `process_source` (:468-496) for every target at once:
    the target starts from generic_init_source, its neighbours from catalog_init_source and stay frozen
    (`init_sources([1], cat_local)`), and `maximize!` runs for all targets of this rank in lock-step on the GPU
    (elbo_maximize.BatchMaximizer).  With world > 1 the targets are sharded by cost; no communication is
    needed because single inference never reads another target's result.  Indices are 0-based."""
    from . import deterministic_vi as dvi
    from .elbo_maximize import BatchMaximizer
    costs = [estimate_time(patches[s, :]) for s in target_sources]
    mine = [target_sources[i] for i in shard_sources(costs, rank, world)]
    field = field or dvi.DeviceField(images, patches)
    rows, act = _task_rows(mine, neighbor_map)
    plan = dvi.Plan(field, rows, act)
    vps = []
    for s, r in zip(mine, rows):
        vps.append(dvi.generic_init_source(catalog[s].pos))
        vps += [dvi.catalog_init_source(catalog[n - 1]) for n in r[1:]]
    bm = BatchMaximizer(plan, np.concatenate(vps), include_kl=include_kl, max_iters=max_iters)
    res = bm.run()
    out = [OptimizedSource(catalog[s].pos[0], catalog[s].pos[1], res.vp[k], bad_sky(catalog[s], images))
           for k, s in enumerate(mine)]
    return out, res


def assign_components(components: List[List[int]], cost_of: Callable[[int], float], n_ranks: int) -> List[List[List[int]]]:
    """shard_batch keeping the component structure: [rank][component][source]."""
    ccost = [sum(cost_of(s) for s in comp) for comp in components]
    order = sorted(range(len(components)), key=lambda i: -ccost[i])
    assign = load_balance([ccost[i] for i in order], n_ranks)
    return [[components[order[i]] for i in idx] for idx in assign]


def bad_sky(ce, images) -> bool:
    """ParallelRun.bad_sky (:437-461): in the i band, the claimed sky (sky x iota at the source's pixel, in
    electrons) is more than 5 photons below the median of the pixels in a 50-pixel box around the source."""
    from .model import box_around_point, clamp_box
    img = next((im for im in images if im.b == 4), None)
    if img is None:
        return False
    pc = img.wcs.world_to_pix(ce.pos)
    h = max(1, min(int(np.rint(pc[0])), img.H))
    w = max(1, min(int(np.rint(pc[1])), img.W))
    claimed_sky = float(np.asarray(img.sky)[h - 1, w - 1]) * float(img.nelec_per_nmgy[h - 1])
    (h0, h1), (w0, w1) = clamp_box(box_around_point(img.wcs, ce.pos, 50.0), (img.H, img.W))
    px = np.asarray(img.pixels)[h0 - 1:h1, w0 - 1:w1]
    px = px[~np.isnan(px)]
    if px.size == 0:
        return False
    return bool(claimed_sky + 5 < float(np.median(px)))


def one_node_joint_infer(catalog, patches, target_sources, neighbor_map, images, field=None, include_kl=True,
                         n_iters=3, batch_size=60, seed=42, max_iters=50, rank=0, world=1, group=None,
                         make_maximizer=None):
    """ParallelRun.one_node_joint_infer (:135-196) + process_sources_dynamic! (:302-370): targets share their
    variational parameters (`ts_vp`, :99-113, 249-252), are visited in Cyclades batches of mutually
    non-conflicting connected components, `num_joint_vi_iters` (= 3, config.jl:20) sweeps.  The reference
    optimises the sources of one component one after another on a thread; here round r optimises the r-th
    source of EVERY component of the batch in one lock-step BatchMaximizer run (sources of different
    components never share pixels, so this is the same serial-equivalent schedule).

    world > 1 (one process per GPU): the components of a batch are dealt to the ranks by cost (a component never
    straddles ranks, so no two GPUs ever touch overlapping patches), every rank optimises its components, and the
    updated parameters are exchanged once per batch -- the reference's thread barrier at the end of a batch
    (:321-324) -- by `allgather_vp` (44 doubles per source).  Every rank returns the full result.

    Each target keeps ONE constraint box for all sweeps: built from its initial parameters like the reference's
    per-target ElboConfig in cfg_vec (:99-101, "configurations must persist so location constraints do not
    shift"), not re-centred on the previous sweep's optimum.

    `make_maximizer(todo, rows, act, vps, box)` -> object with .run() (tests inject a CPU checker); default: the
    CUDA plan + BatchMaximizer.  Indices are 0-based."""
    import torch
    from . import constraint_transforms as ct
    from . import deterministic_vi as dvi
    from .elbo_maximize import BatchMaximizer
    tset = set(target_sources)
    ts_vp = {s: dvi.generic_init_source(catalog[s].pos) for s in target_sources}           # setup_vecs :99-113
    init = torch.as_tensor(np.stack([ts_vp[s] for s in target_sources]), dtype=torch.float64)
    lo_all, hi_all = ct.box_bounds(init)                                                    # one box per target, kept
    box_of = {s: (lo_all[k].numpy().copy(), hi_all[k].numpy().copy()) for k, s in enumerate(target_sources)}
    frozen = {}

    def vp_of(s):
        if s in tset:
            return ts_vp[s]
        if s not in frozen:
            frozen[s] = dvi.catalog_init_source(catalog[s])
        return frozen[s]
    nmap = {s: [n for n in neighbor_map[s]] for s in target_sources}
    batches = partition_cyclades_dynamic(list(target_sources), {s: [n for n in nmap[s] if n in tset] for s in nmap},
                                         batch_size=batch_size, seed=seed)
    cost_cache = {}

    def cost_of(s):
        if s not in cost_cache:
            cost_cache[s] = estimate_time(patches[s, :])
        return cost_cache[s]
    plans = {}
    if make_maximizer is None:
        field = field or dvi.DeviceField(images, patches)

        def make_maximizer(todo, rows, act, vps, box):
            if todo not in plans:
                plans[todo] = dvi.Plan(field, rows, act)
            return BatchMaximizer(plans[todo], vps, include_kl=include_kl, max_iters=max_iters, box=box)
    stats = []
    for _ in range(n_iters):
        for comps in batches:
            mine = assign_components(comps, cost_of, world)[rank] if world > 1 else comps
            for r in range(max((len(c) for c in mine), default=0)):
                todo = tuple(c[r] for c in mine if len(c) > r)
                rows, act = _task_rows(todo, neighbor_map)
                vps = np.concatenate([vp_of(n - 1) for rr in rows for n in rr])
                box = (np.stack([box_of[s][0] for s in todo]), np.stack([box_of[s][1] for s in todo]))
                res = make_maximizer(todo, rows, act, vps, box).run()
                for k, s in enumerate(todo):
                    ts_vp[s][:] = res.vp[k]
                stats.append(res)
            if world > 1:
                # batch barrier (:321-324): everyone learns the parameters the other ranks just optimised
                ids_mine = [s for c in mine for s in c]
                local = np.stack([ts_vp[s] for s in ids_mine], axis=1) if ids_mine else np.zeros((44, 0))
                table = allgather_vp(ids_mine, local, len(catalog), group=group)
                for c in comps:
                    for s in c:
                        ts_vp[s][:] = table[:, s]
    return [OptimizedSource(catalog[s].pos[0], catalog[s].pos[1], ts_vp[s], bad_sky(catalog[s], images))
            for s in target_sources], stats
