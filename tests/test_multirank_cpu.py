"""N > 1 on CPU (gloo, world_size 2): the hot path shards by batch with NO data-path collective - every rank
runs its own batch shard, and the only cross-rank traffic is the max-over-ranks timing reduce of bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

import oracle

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(batch, seed):
    g = np.random.default_rng(seed)
    D, L, N = 6, 96, 16
    f = lambda *s: g.standard_normal(s).astype(np.float32)
    return dict(u=f(batch, D, L), delta=(0.5 * g.random((batch, D, L))).astype(np.float32), B=f(batch, 1, N, L), C=f(batch, 1, N, L),
                z=f(batch, D, L), dout=f(batch, D, L))


def _weights():
    g = np.random.default_rng(123)
    D, N = 6, 16
    return dict(A=(-0.5 * g.random((D, N))).astype(np.float32), D=g.standard_normal(D).astype(np.float32),
                bias=(0.5 * g.random(D)).astype(np.float32))


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    w = _weights()
    t = _inputs(2, seed=rank)                                   # per-rank batch shard, as bench.py seeds by rank
    out, last = oracle.selective_scan_fwd(t["u"], t["delta"], w["A"], t["B"], t["C"], w["D"], t["z"], w["bias"], True)
    g = oracle.selective_scan_bwd(t["u"], t["delta"], w["A"], t["B"], t["C"], w["D"], t["z"], w["bias"], t["dout"], True)
    # weight gradients are the only quantities a data-parallel trainer would all-reduce (outside the hot path)
    dA = torch.from_numpy(np.ascontiguousarray(g["dA"]).astype(np.float64))
    dist.all_reduce(dA)
    gathered = [torch.zeros(out.shape, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(np.ascontiguousarray(out).astype(np.float32)))
    ms = bench.reduce_max_ms(10.0 + 5.0 * rank, dist)            # max over ranks
    if rank == 0:
        np.savez(os.path.join(tmp, "r0.npz"), out=torch.cat(gathered).numpy(), dA=dA.numpy(), ms=ms,
                 value=bench.whole_job_gbps(world, 1e9, ms))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_by_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "r0.npz")
    w = _weights()
    shards = [_inputs(2, seed=k) for k in range(world)]
    full = {k: np.concatenate([s[k] for s in shards]) for k in shards[0]}
    out, _ = oracle.selective_scan_fwd(full["u"], full["delta"], w["A"], full["B"], full["C"], w["D"], full["z"], w["bias"], True)
    g = oracle.selective_scan_bwd(full["u"], full["delta"], w["A"], full["B"], full["C"], w["D"], full["z"], w["bias"], full["dout"], True)
    assert np.array_equal(r["out"], out.astype(np.float32))       # bit-exact: a batch element never sees another rank's data
    np.testing.assert_allclose(r["dA"], g["dA"], rtol=1e-5, atol=1e-5)    # sum of per-rank weight grads == full-batch grad
    assert float(r["ms"]) == 15.0                                 # max over ranks
    assert abs(float(r["value"]) - 2 * 1e9 / 15e-3 / 1e9) < 1e-9  # whole-job value = all ranks' bytes / max time


# ---- the caller of the path: data-parallel MM_Net training, gloo, world_size 2 ---------------------------
def _ddp_train_worker(rank, world, port, q):
    try:
        _ddp_train_body(rank, world, port, q)
    except Exception as exc:   # report instead of leaving the parent waiting on the queue
        q.put((rank, repr(exc)))
        raise


def _ddp_train_body(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.join(ROOT, "mm-unet_b200"))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import torch_ref
    from mmunet_b200 import mamba as mamba_mod, mm_net, train

    class OracleMamba(mamba_mod.Mamba):          # CPU stand-in for the CUDA Mamba block (tests only)
        def forward(self, hidden_states, inference_params=None):
            return torch_ref.mamba_forward(self, hidden_states)

    mm_net.Mamba, mm_net._flatten_two_row, mm_net._unflatten_two_row = OracleMamba, torch_ref.two_row_flatten, torch_ref.two_row_unflatten
    mm_net._snake_sample = None
    tr = train.Trainer(image_size=64, batch_per_rank=2, dtype="fp32", device="cpu")
    tr.set_epoch(2)
    x, y = tr.synthetic_batch()
    w0 = tr.net.encoder1[0].weight.detach().clone()
    losses = [float(tr.step(x, y)) for _ in range(2)]
    # after the all-reduced update every rank must hold the same weights although the ranks saw different batches
    w = tr.net.encoder1[0].weight.detach().clone()
    gathered = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    frozen = [n for n, p in tr.net.named_parameters() if not p.requires_grad]
    q.put((rank, losses, bool(torch.equal(gathered[0], gathered[1])), float((w - w0).abs().max()), len(frozen),
           float(x.sum())))
    dist.destroy_process_group()


def test_ddp_trainer_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(len(r) > 2 for r in res), res
    (r0, l0, same0, moved0, nfrozen0, xs0), (r1, l1, same1, moved1, nfrozen1, xs1) = res
    assert same0 and same1, "ranks diverged: the gradient all-reduce did not average the update"
    assert moved0 > 0 and np.isfinite(l0 + l1).all()
    assert xs0 != xs1, "ranks must train on different shards"
    assert nfrozen0 == nfrozen1 == 47 * (2 + 14), "dsc_conv_y (2) + the _b/_s Mamba sets (2 x 7 tensors) of the 47 MMConvs are frozen"


def test_reference_arm_uses_all_cores_under_torchrun_env(monkeypatch, capfd):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm (rank 0 only) must still run the C oracle on every
    core the process may use, and the other ranks must print nothing.  Shapes shrunk so the test takes seconds."""
    import argparse, json, os, sys
    sys.path.insert(0, ROOT)
    import bench, oracle
    monkeypatch.setattr(bench, "B", 2); monkeypatch.setattr(bench, "D", 16); monkeypatch.setattr(bench, "L", 256)
    args = argparse.Namespace(gpus=2, steps=1, warmup=0)
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    oracle.set_num_threads(1)                                   # what a worker started by torchrun sees
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args)
    assert capfd.readouterr().out.strip() == ""
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(args)
    line = json.loads(capfd.readouterr().out.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == cores and line["n_gpus"] == 2
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
