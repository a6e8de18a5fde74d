"""NCCL-side correctness of the data-parallel path (VERDICT r1 missing #7): with the global batch
split into equal shards, the all-reduced flat gradient bucket on every rank equals the mean of the per-shard
gradients computed on one GPU.  Needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`
(skips on a one-GPU box; the gloo twin runs on CPU in tests/test_distributed.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model(device):
    from cgc_net_b200 import network
    torch.manual_seed(0)
    return network.SoftPoolingGcnEncoder(3000, 16, 20, 20, True, True, 20, 3, 0.1, [50], concat=True,
                                         load_data_sparse=True).to(device).train()


def _graphs(world):
    from cgc_net_b200 import graph
    return [graph.synthetic_cell_graph(600, 8, 16, seed=50 + i) for i in range(4 * world)]


def _shard_grads(device, shard, overlap=0):
    from cgc_net_b200 import graph, parallel
    model = _model(device)
    bucket = parallel.FlatGradBucket(model.parameters())
    if overlap:
        assert bucket.enable_overlap(chunks=overlap)          # chunk all-reduces start inside backward()
    bucket.zero()
    _, loss = model(graph.Batch.from_data_list(shard).to(device))
    loss.backward()
    return model, bucket


def _worker(rank, world, port, out, overlap=0):
    from cgc_net_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    device = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=device)
    graphs = _graphs(world)
    model, bucket = _shard_grads(device, parallel.shard_for_rank(graphs, rank, world), overlap)
    bucket.all_reduce_mean()
    torch.cuda.synchronize()
    flat = bucket.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        same = all(torch.equal(gathered[0], g) for g in gathered)         # every rank holds the same reduced bucket
        # single-GPU truth: mean over the shards of each shard's own gradient, all computed on this GPU
        want = torch.zeros_like(flat)
        for r in range(world):
            _, b = _shard_grads(device, parallel.shard_for_rank(graphs, r, world))
            b.pack()
            want += b.flat
        want /= world
        torch.save({'same': same, 'got': flat.cpu(), 'want': want.cpu()}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('overlap', [0, 2])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_nccl_allreduce_matches_single_gpu(cuda_backend, tmp_path, world, overlap):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    out = str(tmp_path / 'rank0.pt')
    mp.spawn(_worker, args=(world, _free_port(), out, overlap), nprocs=world, join=True)
    res = torch.load(out)
    got, want = res['got'], res['want']
    diff = float((got - want).abs().max())
    print('world %d overlap %d: bucket of %d gradients, max |allreduced - single-GPU mean| = %.3e (scale %.3e), bitwise %s'
          % (world, overlap, got.numel(), diff, float(want.abs().max()), bool(torch.equal(got, want))))
    assert res['same'], 'ranks disagree on the reduced bucket'
    assert diff <= 1e-6 * float(want.abs().max()) + 1e-9        # fp32 reduction order only
