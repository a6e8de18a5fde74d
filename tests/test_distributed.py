"""Data-parallel path on CPU: world_size 2 over gloo (host logic of cgc_net_b200/parallel.py).

Two ranks each take a shard of the graphs, run forward/backward (per-operator CPU formulas injected as the
backend), all-reduce ONE flat gradient bucket and must end up with the mean of the per-replica gradients --
the reference's DataParallel semantics (mean of per-replica mean losses, train.py:179; per-replica BatchNorm)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from cgc_net_b200 import _backend, graph, network, parallel
from oracle import op_reference


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_model():
    torch.manual_seed(3)
    return network.SoftPoolingGcnEncoder(*helpers.encoder_args(300), concat=True, load_data_sparse=True).train()


def _graphs():
    return helpers.make_graphs([40, 55, 33, 61], [5, 6, 4, 8], seed=9)


def _worker(rank, world, port, out, overlap=0):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    _backend.use_backend_for_testing(op_reference.TorchBackend())
    model = _make_model()
    if rank != 0:                                     # ranks start different, rank 0's weights win
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    parallel.broadcast_parameters(model, src=0)
    bucket = parallel.FlatGradBucket(model.parameters())
    launched = []
    if overlap:
        assert bucket.enable_overlap(chunks=overlap)
        orig = bucket._overlap._launch
        bucket._overlap._launch = lambda c: (launched.append(c), orig(c))[1]
    shard = parallel.shard_for_rank(_graphs(), rank, world)
    for _ in range(2 if overlap else 1):              # overlapped mode: a second step checks the per-step reset
        bucket.zero()
        _, loss = model(graph.Batch.from_data_list(shard))
        loss.backward()
        early = list(launched)                         # chunks whose all-reduce started inside backward()
        bucket.all_reduce_mean()
        launched.clear()
    if rank == 0:
        torch.save({'flat': bucket.flat.clone(), 'loss': loss.detach(), 'early': early,
                    'grads_are_views': all(p.grad.data_ptr() == v.data_ptr()
                                           for p, v in zip(bucket.params, bucket.views))}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize('overlap', [0, 1, 3])
def test_two_rank_gradient_allreduce(tmp_path, torch_backend, overlap):
    """overlap = 0: one all-reduce after backward; 1 / 3: FlatGradBucket.enable_overlap with that many chunks (the
    chunk all-reduces start from post-accumulate-grad hooks inside backward()) -- same reduced bucket."""
    out = str(tmp_path / 'rank0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out, overlap), nprocs=2, join=True)
    got = torch.load(out)
    assert got['grads_are_views']
    if overlap:
        assert len(got['early']) >= 1, 'no chunk was reduced inside backward()'
    # single-process reference: mean over the two shards of each shard's own gradient
    graphs = _graphs()
    flats = []
    for rank in range(2):
        model = _make_model()
        bucket = parallel.FlatGradBucket(model.parameters())
        bucket.zero()
        _, loss = model(graph.Batch.from_data_list(parallel.shard_for_rank(graphs, rank, 2)))
        loss.backward()
        bucket.pack()
        flats.append(bucket.flat.clone())
    want = (flats[0] + flats[1]) / 2
    assert got['flat'].numel() == sum(p.numel() for p in _make_model().parameters())
    torch.testing.assert_close(got['flat'], want, rtol=1e-5, atol=1e-7)


def test_flat_bucket_views_and_sharding():
    model = _make_model()
    bucket = parallel.FlatGradBucket(model.parameters())
    bucket.zero()
    assert all(p.grad is None for p in model.parameters())      # autograd will assign, not accumulate
    for i, p in enumerate(model.parameters()):
        if i % 2 == 0:                                           # half of the gradients present, half missing
            p.grad = torch.full_like(p, 2.0)
    bucket.pack()
    lo = bucket.flat.data_ptr()
    hi = lo + bucket.nbytes()
    assert all(lo <= p.grad.data_ptr() < hi for p in model.parameters())        # views of the one buffer
    for i, p in enumerate(model.parameters()):
        assert float(p.grad.min()) == float(p.grad.max()) == (2.0 if i % 2 == 0 else 0.0)
    bucket.flat.fill_(3.0)
    assert all(float(p.grad.min()) == 3.0 for p in model.parameters())
    bucket.all_reduce_mean()                           # no process group: a no-op
    graphs = _graphs()
    assert sum(len(parallel.shard_for_rank(graphs, r, 2)) for r in range(2)) == len(graphs)
    by_nodes = [parallel.shard_for_rank(graphs, r, 2, by_nodes=True) for r in range(2)]
    assert sum(len(s) for s in by_nodes) == len(graphs)


def test_flat_bucket_pack_is_idempotent():
    """pack() twice, pack() + all_reduce_mean(), and accumulation into the live views must keep the gradients
    (the second pack() used to zero the whole buffer)."""
    model = _make_model()
    bucket = parallel.FlatGradBucket(model.parameters())
    bucket.zero()
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(i + 1))
    bucket.pack()
    first = bucket.flat.clone()
    assert float(first.abs().min()) > 0
    bucket.pack()                                       # every .grad already is its view
    assert torch.equal(bucket.flat, first)
    bucket.all_reduce_mean()                            # no process group: no-op, still intact
    assert torch.equal(bucket.flat, first)
    for p in model.parameters():                        # gradient accumulation without zero(): grads live in the views
        p.grad += 1.0
    bucket.pack()
    assert torch.equal(bucket.flat, first + 1.0)
    params = list(model.parameters())
    params[0].grad = None                               # one gradient missing -> only ITS view is zeroed
    bucket.pack()
    assert float(params[0].grad.abs().max()) == 0.0
    assert torch.equal(params[1].grad, torch.full_like(params[1], 3.0))
