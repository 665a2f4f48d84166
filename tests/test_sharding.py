"""world_size-2 gloo test of the N>1 host logic: row-shard the pods, run the dense stage per rank (the device
algorithm compiled for the host, tests/native/emul.cpp, stands in for the kernels — no GPU in this container),
all-gather the COMPACT per-pod results (template + attempt codes, 8 bytes a pod — the instance-type rows are not
sent), rebuild them on rank 0 and compare with the unsharded evaluation and with the oracle.  The GPU version of
the same path is bench.py under torchrun (b200s_resident_export_compact / _import_compact / _pack + NCCL)."""
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
    import parity
    from b200sched import capi, model as M, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth.config("c3", 1001, 130)          # odd row count: ranks get unequal shards
    lo, hi = sharding.row_shard(m.n_pods, rank, world)
    enc = capi.Encoded(M.slice_pods(m, lo, hi))
    fit = parity.emul_fit_matrix(enc)          # this rank's rows (they stay here) ...
    compact = sharding.to_compact(fit.tmpl, fit.code)  # ... and what the packing rank needs of them
    rows = max(sharding.row_shard(m.n_pods, r, world)[1] - sharding.row_shard(m.n_pods, r, world)[0] for r in range(world))
    t = torch.from_numpy(sharding.pad_rows(compact, rows))
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)                      # the one collective
    ev = torch.tensor([fit.evals])
    dist.all_reduce(ev)
    if rank == 0:
        enc_full = capi.Encoded(m)
        got_tmpl, got_code = sharding.from_compact(sharding.unpad_gathered(torch.stack(g).numpy(), m.n_pods, world), enc_full.T)
        full = parity.emul_fit_matrix(enc_full)
        cf = parity.canon_fit(enc_full, capi.FitOut(got_tmpl, full.mask, got_code, int(ev.item())))
        o_tmpl, o_mask, o_code, o_evals, _ = ob.fit_matrix(m)
        ok = (np.array_equal(got_tmpl, full.tmpl) and np.array_equal(got_code, full.code) and int(ev.item()) == full.evals and
              np.array_equal(cf["tmpl"], o_tmpl) and np.array_equal(cf["code"], o_code) and int(ev.item()) == o_evals)
        # "merge the global packing": the pack over all pods (rows re-derived from the class tables) equals the oracle's
        raw = parity.emul_solve(enc_full)
        try:
            parity.assert_same(parity.canon_device(enc_full, raw), parity.canon_oracle(m, ob.solve(m)), "global pack")
        except AssertionError:
            ok = False
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_compact_rows_round_trip():
    from b200sched import sharding
    rng = np.random.default_rng(3)
    for T in (1, 4, 7, 8, 16):
        tmpl = rng.integers(-1, T, 50).astype(np.int32)
        code = rng.integers(0, 256, (50, T)).astype(np.uint8)
        rows = sharding.to_compact(tmpl, code)
        assert rows.shape == (50, sharding.compact_row_bytes(T))
        t2, c2 = sharding.from_compact(rows, T)
        assert np.array_equal(t2, tmpl) and np.array_equal(c2, code)


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
