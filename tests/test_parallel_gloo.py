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
        bucket = FlatGradBucket(model.parameters(), n_chunks=2)
        assert bucket.numel == 16 + 16 * 4 + 4 and bucket.n_chunks == 2
        # the buffer is laid out in backward-completion order: the LAST module's parameters first
        assert model[2].bias.grad.data_ptr() == bucket.flat[:4].data_ptr()
        assert model[0].bias.grad.data_ptr() == bucket.flat[-16:].data_ptr()
        data = torch.arange(6 * 8, dtype=torch.float32).reshape(6, 8) / 10
        sl = shard_batch(6, rank, world)
        x = data[sl]
        assert x.shape[0] == 3
        loss = model(x).pow(2).mean()
        loss.backward()  # hooks not armed: autograd accumulates into the zeroed views
        local = bucket.flat.clone()
        bucket.all_reduce_mean_async()
        bucket.wait()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean_grad = sum(gathered) / world
        assert torch.allclose(bucket.flat, mean_grad, atol=1e-7)
        norm = bucket.clip_grad_norm_(1e-3)
        assert torch.linalg.vector_norm(bucket.flat) <= 1e-3 + 1e-6 and norm > 0
        # step protocol: every chunk is gathered AND all-reduced from inside backward (post-accumulate-grad hooks)
        bucket.begin_step()
        assert all(p.grad is None for p in bucket.params)
        n0 = bucket.chunks_reduced_in_backward
        model(x).pow(2).mean().backward()
        assert bucket.chunks_reduced_in_backward - n0 == bucket.n_chunks
        assert model[2].bias.grad.data_ptr() == bucket.flat[:4].data_ptr()
        bucket.finish_backward()
        assert torch.allclose(bucket.flat, mean_grad, atol=1e-7)
        # accum_grad = 2: the first micro-step stays local (no collective), the second adds and reduces the sum
        x2 = x * 0.5 + 0.1
        bucket.begin_step(accumulate=False, sync=False)
        n0 = bucket.chunks_reduced_in_backward
        model(x).pow(2).mean().backward()
        bucket.finish_backward()
        assert bucket.chunks_reduced_in_backward == n0 and torch.allclose(bucket.flat, local, atol=1e-7)
        bucket.begin_step(accumulate=True, sync=True)
        model(x2).pow(2).mean().backward()
        bucket.finish_backward()
        bucket.zero_()
        model(x2).pow(2).mean().backward()
        local2 = bucket.flat.clone()
        g2 = [torch.zeros_like(local2) for _ in range(world)]
        dist.all_gather(g2, local2)
        # (the buffer was zeroed for the check above; redo the accumulation to compare)
        bucket.begin_step(accumulate=False, sync=False)
        model(x).pow(2).mean().backward()
        bucket.finish_backward()
        bucket.begin_step(accumulate=True, sync=True)
        model(x2).pow(2).mean().backward()
        bucket.finish_backward()
        assert torch.allclose(bucket.flat, mean_grad + sum(g2) / world, atol=1e-6)
        # trainer.py:677 — a non-finite gradient norm skips the update (and is counted)
        from aga_b200.graphed import EagerTrainStep, GuardedUpdate
        opt = torch.optim.SGD(bucket.params, lr=0.1)
        upd = GuardedUpdate(opt, bucket, 1.0)
        before = [p.detach().clone() for p in bucket.params]
        bucket.flat[3] = float("nan")
        upd()
        assert all(torch.equal(a, p.detach()) for a, p in zip(before, bucket.params)) and float(upd.skipped_steps) == 1.0
        bucket.flat.copy_(mean_grad)
        upd()
        assert not torch.equal(before[0], bucket.params[0].detach()) and float(upd.skipped_steps) == 1.0
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
