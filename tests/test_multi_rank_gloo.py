"""N > 1 path on CPU: two gloo ranks shard the chains of one run_mcmc() call, run their slices and
gather on rank 0.  The CPU oracle stands in for the CUDA engine (the test replaces engine.run_raw),
so what is under test is the product's sharding, chain_offset bookkeeping and gather; the result must
equal the unsharded run bit for bit (chains never interact; the Philox counter carries the global
chain id)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import drjacoby_b200 as drj
    from drjacoby_b200 import api
    from oracle import pyoracle
    from test_host_api import oracle_runner
    api.engine.run_raw = oracle_runner(pyoracle)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(1).normal(3, 2, 30)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.use_builtin("example")
    res = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=30, samples=40, rungs=3, chains=5,
                       seed=11, cluster="torch", silent=True)
    if rank == 0:
        np.savez(os.path.join(tmpdir, "sharded.npz"), **res.raw, rhat=res.diagnostics["rhat"].to_numpy())
    else:
        assert res is None
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, oracle):
    import torch.multiprocessing as mp
    import drjacoby_b200 as drj
    from drjacoby_b200 import api
    from test_host_api import oracle_runner
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    got = np.load(os.path.join(tmp_path, "sharded.npz"))
    old = api.engine.run_raw
    api.engine.run_raw = oracle_runner(oracle)
    try:
        x = np.random.default_rng(1).normal(3, 2, 30)
        df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
        drj.use_builtin("example")
        ref = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=30, samples=40, rungs=3, chains=5,
                           seed=11, silent=True)
    finally:
        api.engine.run_raw = old
    for k, v in ref.raw.items():
        if k != "t_diff":
            assert np.array_equal(got[k], v), k
    assert np.array_equal(got["rhat"], ref.diagnostics["rhat"].to_numpy())


def _sparse_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import drjacoby_b200 as drj
    from drjacoby_b200 import api
    from oracle import pyoracle
    from test_host_api import oracle_runner
    api.engine.run_raw = oracle_runner(pyoracle)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(1).normal(3, 2, 30)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.use_builtin("example")
    res = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=20, samples=20, rungs=2, chains=2,
                       seed=11, cluster="torch", silent=True, save_hot_draws=True)
    if rank == 0:
        np.savez(os.path.join(tmpdir, "sparse.npz"), **res.raw)
    else:
        assert res is None
    dist.destroy_process_group()


def test_more_ranks_than_chains(tmp_path, oracle):
    """The reference's default is 5 chains; on an 8-GPU box ranks 5..7 have no chains.  They must skip the native
    call and still enter every collective (they used to fail before the gather and leave the others hanging)."""
    import torch.multiprocessing as mp
    import drjacoby_b200 as drj
    from drjacoby_b200 import api
    from test_host_api import oracle_runner
    port = 33500 + (os.getpid() % 2000)
    mp.start_processes(_sparse_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True, start_method="spawn")
    got = np.load(os.path.join(tmp_path, "sparse.npz"))
    old = api.engine.run_raw
    api.engine.run_raw = oracle_runner(oracle)
    try:
        x = np.random.default_rng(1).normal(3, 2, 30)
        df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
        drj.use_builtin("example")
        ref = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=20, samples=20, rungs=2, chains=2,
                           seed=11, silent=True, save_hot_draws=True)
    finally:
        api.engine.run_raw = old
    for k, v in ref.raw.items():
        if k != "t_diff":
            assert got[k].shape == v.shape and np.array_equal(got[k], v), k


def _gather_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from drjacoby_b200 import api
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api._PACK_LIMIT = 1000  # bytes: "big" travels on its own, the others in the packed message
    n = 2 + rank            # uneven shards
    rng = np.random.default_rng(rank)
    raw = {"big": rng.normal(size=(n, 70, 3)), "counts": rng.integers(0, 9, size=(n, 4)).astype(np.int32), "empty": np.zeros((n, 0), np.int32),
           "t": rng.normal(size=n)}
    out = api.gather_shards(raw, 2 + 3 + 4, rank, world)
    np.savez(os.path.join(tmpdir, f"in{rank}.npz"), **raw)
    if rank == 0:
        np.savez(os.path.join(tmpdir, "out.npz"), **out)
    else:
        assert out is None
    dist.destroy_process_group()


def test_gather_shards_packed_and_large_paths(tmp_path):
    """gather_shards: small arrays of all dtypes in one packed message, large ones one by one, uneven shard sizes,
    zero-width arrays; rank 0 gets the concatenation in rank order."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.start_processes(_gather_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True, start_method="spawn")
    out = np.load(os.path.join(tmp_path, "out.npz"))
    ins = [np.load(os.path.join(tmp_path, f"in{r}.npz")) for r in range(3)]
    for k in ("big", "counts", "empty", "t"):
        want = np.concatenate([i[k] for i in ins], axis=0)
        assert out[k].dtype == want.dtype and np.array_equal(out[k], want), k
