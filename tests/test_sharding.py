"""world_size-2 gloo test of the N>1 host logic: row-shard the pods, run the dense stage per rank (the CPU oracle
stands in for the kernel here — no GPU in this container), all-gather the result rows, compare with the unsharded
matrix.  The GPU version of the same path is bench.py --workload c5 under torchrun."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
    import oracle_binding as ob
    from b200sched import model as M, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth.config("c3", 1001, 130)          # odd row count: ranks get unequal shards
    lo, hi = sharding.row_shard(m.n_pods, rank, world)
    shard = M.slice_pods(m, lo, hi)
    tmpl, mask, code, evals, T = ob.fit_matrix(shard)
    rows = max(sharding.row_shard(m.n_pods, r, world)[1] - sharding.row_shard(m.n_pods, r, world)[0] for r in range(world))
    outs = []
    for a in (tmpl.reshape(-1, 1).astype(np.int64), mask.astype(np.int64), code.astype(np.int64)):
        t = torch.from_numpy(sharding.pad_rows(a, rows))
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        outs.append(sharding.unpad_gathered(torch.stack(g).numpy(), m.n_pods, world))
    ev = torch.tensor([evals])
    dist.all_reduce(ev)
    if rank == 0:
        f_tmpl, f_mask, f_code, f_evals, _ = ob.fit_matrix(m)
        ok = (np.array_equal(outs[0][:, 0], f_tmpl) and np.array_equal(outs[1].astype(np.uint64), f_mask) and
              np.array_equal(outs[2].astype(np.uint8), f_code) and int(ev.item()) == f_evals)
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_row_shard_is_a_partition():
    from b200sched import sharding
    for n in (0, 1, 7, 8, 1001):
        for w in (1, 2, 3, 8):
            spans = [sharding.row_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_dense_stage_with_gloo_allgather(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok")).read() == "1"


def _worker_deal(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
    import oracle_binding as ob
    import parity
    from b200sched import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model, node_of_pod, util = synth.c4(1200, 100, solve=ob.solve)
    sims = synth.c4_simulations(model, node_of_pod, util, 11)      # odd count: unequal shares
    prices, spot = synth.candidate_prices(model, [i for i, _ in sims])
    _, v_spot, v_od, v_res = parity.capacity_type_values(model)

    def command(j):  # the CPU oracle stands in for solve_batch + consolidation_commands
        out, cmd = ob.consolidate(sims[j][1], prices[j], spot[j], 1, v_spot, v_od, v_res, True)
        return [cmd["decision"], cmd["reason"], cmd["n_options"], out.n_claims, out.n_unscheduled]

    mine = sharding.deal(len(sims), rank, world)
    rows = sharding.dealt_rows(len(sims), world)
    t = torch.from_numpy(sharding.pad_rows(np.array([command(int(j)) for j in mine], np.int64).reshape(-1, 5), rows))
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    got = sharding.undeal(torch.stack(g).numpy(), len(sims), world)
    if rank == 0:
        want = np.array([command(j) for j in range(len(sims))], np.int64)
        open(os.path.join(out_dir, "ok"), "w").write("1" if np.array_equal(got, want) else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_deal_is_a_partition():
    from b200sched import sharding
    for n in (0, 1, 2, 7, 256, 257):
        for world in (1, 2, 3, 8):
            shares = [sharding.deal(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(shares).tolist()) == list(range(n))
            assert max(len(s) for s in shares) <= sharding.dealt_rows(n, world) if n else True
            g = np.zeros((world, max(sharding.dealt_rows(n, world), 1), 2), np.int64)
            for r, idx in enumerate(shares):
                g[r, :len(idx), 0] = idx
                g[r, :len(idx), 1] = idx * 3
            back = sharding.undeal(g, n, world)
            assert back[:, 0].tolist() == list(range(n)) and back[:, 1].tolist() == [3 * i for i in range(n)]


def test_consolidation_simulations_dealt_across_ranks_with_gloo(tmp_path):
    """config 4 at N > 1: simulations are independent, so ranks take them round-robin and all-gather five words per
    simulation.  bench.py --workload c4 --gpus N is the same flow with the device path per rank and NCCL."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_deal, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok")).read() == "1"
