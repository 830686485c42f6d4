"""Partitioning / sharding host logic (src/partition.jl, ParallelRun.jl:28-95) and the N > 1 path over gloo
with world_size = 2 on CPU.  The evaluator handed to `evaluate_sharded` here is the oracle (checker); in
production it is the CUDA plan -- the sharding code is the same."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import oracle_lib
from celeste_jl_b200 import parallel_run as pr
from celeste_jl_b200 import synthetic
from celeste_jl_b200.model import find_all_neighbors, find_neighbors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def field():
    return synthetic.FieldDataset(120, H=300, W=260, seed=5, device="cpu")


def test_find_neighbors_matches_vectorised(field):
    nb = find_all_neighbors(field.patches)
    for t in (0, 17, 63, 119):
        assert find_neighbors(field.patches, t) == nb[t]
    for t, l in enumerate(nb):
        for s in l:
            assert t in nb[s]          # box overlap is symmetric


def test_cyclades_covers_all_sources_without_conflicts(field):
    """test/test_partition.jl:20-94: every source appears exactly once; within a batch no conflict edge joins
    two different components (so ranks never share pixels); repeated with several seeds."""
    nmap = {s: field.neighbors[s] for s in range(len(field.catalog))}
    for seed in range(20):
        batches = pr.partition_cyclades_dynamic(list(nmap), nmap, batch_size=30, seed=seed)
        seen = [s for b in batches for c in b for s in c]
        assert sorted(seen) == list(range(len(nmap)))
        for b in batches:
            comp_of = {s: k for k, c in enumerate(b) for s in c}
            for s, k in comp_of.items():
                for nb in nmap[s]:
                    if nb in comp_of:
                        assert comp_of[nb] == k
            shards = pr.shard_batch(b, lambda s: pr.estimate_time(field.patches[s, :]), 4)
            assert sorted(s for sh in shards for s in sh) == sorted(comp_of)
            rank_of = {s: r for r, sh in enumerate(shards) for s in sh}
            for s in comp_of:
                for nb in nmap[s]:
                    if nb in rank_of:
                        assert rank_of[nb] == rank_of[s]


def test_partition_equally():
    parts = pr.partition_equally(3, 10)
    assert [len(p) for p in parts] == [3, 3, 4] and sum(parts, []) == list(range(10))


def test_shard_sources_is_a_balanced_partition(field):
    costs = [pr.estimate_time(field.patches[s, :]) for s in range(len(field.catalog))]
    for world in (1, 2, 4, 8):
        shards = [pr.shard_sources(costs, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(len(costs)))
        load = [sum(costs[s] for s in sh) for sh in shards]
        assert max(load) - min(load) <= max(costs)


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
import oracle_lib
from celeste_jl_b200 import parallel_run as pr, synthetic
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ds = synthetic.FieldDataset(60, H=220, W=200, seed=9, device="cpu")
costs = [pr.estimate_time(ds.patches[s, :]) for s in range(len(ds.catalog))]
of = oracle_lib.OracleField(ds.images, ds.patches)
def evaluate(idx):
    rows, act = ds.tasks(idx)
    return of.elbo_batch([(r, a, np.stack([ds.vp[i - 1] for i in r], axis=1)) for r, a in zip(rows, act)], mode=1)
mine, out, total = pr.evaluate_sharded(len(costs), costs, evaluate, rank, world)
vp_local = np.stack([ds.vp[s] + rank + 1 for s in mine], axis=1)
table = pr.allgather_vp(mine, vp_local, len(costs))
if rank == 0:
    full = evaluate(list(range(len(costs))))
    owner = {s: r for r in range(world) for s in pr.shard_sources(costs, r, world)}
    ok_vp = all(np.array_equal(table[:, s], ds.vp[s] + owner[s] + 1) for s in range(len(costs)))
    print(json.dumps({"total": total, "full": float(full["v"].sum()), "n_mine": len(mine), "ok_vp": bool(ok_vp)}))
dist.destroy_process_group()
'''


def test_two_rank_gloo_shard_and_reduce(tmp_path):
    """N > 1 path on CPU: two ranks shard the sources, each evaluates its shard, the ELBO all-reduce equals the
    single-process sum and the vp all-gather delivers every owner's update."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok_vp"] and 0 < res["n_mine"] < 60
    assert abs(res["total"] - res["full"]) <= 1e-12 * abs(res["full"])


def test_bad_sky_flags_underestimated_background():
    """ParallelRun.jl:437-461: claimed sky (electrons) + 5 < median of the 50-pixel box in the i band."""
    ds = synthetic.FieldDataset(3, H=140, W=120, seed=2, device="cpu")
    assert not any(pr.bad_sky(ce, ds.images) for ce in ds.catalog)
    img = next(im for im in ds.images if im.b == 4)
    img.pixels = img.pixels + np.float32(40.0)          # 40 unexplained electrons per pixel
    assert all(pr.bad_sky(ce, ds.images) for ce in ds.catalog)
    assert not pr.bad_sky(ds.catalog[0], [im for im in ds.images if im.b != 4])


JOINT_WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from celeste_jl_b200 import parallel_run as pr, synthetic, elbo_maximize as em
from test_maximize import OracleRunner, PlanLike
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("gloo")
rank = dist.get_rank() if world > 1 else 0
ds = synthetic.FieldDataset(14, H=70, W=64, seed=31, device="cpu")
nmap = {s: ds.neighbors[s] for s in range(len(ds.catalog))}
boxes = []
def make(todo, rows, act, vps, box):
    pl = PlanLike(rows, act)
    boxes.append((todo, box[0][:, :2].copy()))
    return em.BatchMaximizer(pl, vps, include_kl=True, device="cpu", max_iters=4, box=box,
                             runner=OracleRunner(ds.images, ds.patches, pl))
res, stats = pr.one_node_joint_infer(ds.catalog, ds.patches, list(range(14)), nmap, ds.images, n_iters=2, batch_size=7,
                                     max_iters=4, rank=rank, world=world, make_maximizer=make)
first = {}
same_box = True
for todo, lo in boxes:
    for k, s in enumerate(todo):
        if s in first:
            same_box &= bool(np.array_equal(first[s], lo[k]))
        else:
            first[s] = lo[k]
if rank == 0:
    print(json.dumps({"vp": np.stack([r.vs for r in res]).tolist(), "same_box": same_box, "visited": len(first),
                      "runs": len(stats)}))
if world > 1:
    dist.destroy_process_group()
'''


def test_joint_infer_two_ranks_equals_one_rank_and_keeps_its_boxes(tmp_path):
    """one_node_joint_infer with rank / world (components of a Cyclades batch dealt to ranks, vp all-gather at the
    batch barrier, ParallelRun.jl:321-324) gives the same catalog as the single-process run; and every target is
    optimised inside the SAME position box on every sweep (ParallelRun.jl:99-101).  Evaluator: the oracle."""
    import json
    script = tmp_path / "joint_worker.py"
    script.write_text(JOINT_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    one = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, env=env, timeout=900)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=900)
    assert two.returncode == 0, two.stderr[-2000:]
    a = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    b = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    assert a["same_box"] and b["same_box"] and a["visited"] == 14
    assert b["runs"] <= a["runs"]
    assert np.allclose(np.array(a["vp"]), np.array(b["vp"]), rtol=1e-9, atol=1e-12)
