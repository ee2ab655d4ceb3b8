"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: flat gradient bucket all-reduce == DDP averaging,
packed stats all-reduce, batch[rank::world] sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import aga_b200  # noqa: F401
        from aga_b200.parallel import FlatGradBucket, all_reduce_stats, shard_batch

        torch.manual_seed(0)  # same weights on every rank
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
        model[0].weight.requires_grad_(False)  # frozen params are not part of the bucket
        bucket = FlatGradBucket(model.parameters())
        assert bucket.numel == 16 + 16 * 4 + 4
        data = torch.arange(6 * 8, dtype=torch.float32).reshape(6, 8) / 10
        sl = shard_batch(6, rank, world)
        x = data[sl]
        assert x.shape[0] == 3
        loss = model(x).pow(2).mean()
        loss.backward()
        local = bucket.flat.clone()
        bucket.all_reduce_mean_async()
        bucket.wait()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat, sum(gathered) / world, atol=1e-7)
        # grads are views into the flat buffer
        assert model[2].bias.grad.data_ptr() == bucket.flat[-4:].data_ptr()
        norm = bucket.clip_grad_norm_(1e-3)
        assert torch.linalg.vector_norm(bucket.flat) <= 1e-3 + 1e-6 and norm > 0
        # assign-then-gather step protocol: same flat contents as accumulating into zeroed views
        bucket.zero_()
        model(x).pow(2).mean().backward()
        ref = bucket.flat.clone()
        bucket.begin_step()
        assert all(p.grad is None for p in bucket.params)
        model(x).pow(2).mean().backward()
        assert model[2].bias.grad.data_ptr() != bucket.flat[-4:].data_ptr()
        bucket.gather_()
        assert torch.equal(bucket.flat, ref) and model[2].bias.grad.data_ptr() == bucket.flat[-4:].data_ptr()
        stats = all_reduce_stats({"loss": loss.detach(), "acc": torch.tensor(float(rank)), "cer": None},
                                 torch.tensor(float(x.shape[0])))
        assert stats["cer"] is None and abs(float(stats["acc"]) - 0.5) < 1e-6
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
