"""A/B of Python-side switches of the training step at full size (B = 64 x 1800, config #3): step time (fwd + bwd, CUDA
events, 5 repetitions after a warm-up) and the largest gradient difference against the first arm.
Usage (GPU box): python tools/ab_step.py FOLD_BN_BACKWARD=0 FOLD_BN_BACKWARD=1 [...]
Each argument is one arm: comma-separated NAME=VALUE pairs applied to module constants of cgc_net_b200.ops."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cgc_net_b200 import graph, network, ops  # noqa: E402

dev = torch.device('cuda:0')
batch = graph.Batch.from_data_list(graph.synthetic_batch(64, 1800, 8, 16, seed=5)).to(dev)
torch.manual_seed(1)
model = network.SoftPoolingGcnEncoder(11404, 16, 20, 20, True, True, 20, 3, 0.1, [50], concat=True,
                                      load_data_sparse=True).to(dev).train()


def step():
    model.zero_grad(set_to_none=True)
    _, loss = model(batch)
    loss.backward()


def timed(reps=5):
    step()
    step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


first = None
for arm in sys.argv[1:] or ['FOLD_BN_BACKWARD=0', 'FOLD_BN_BACKWARD=1']:
    for kv in arm.split(','):
        k, v = kv.split('=')
        assert hasattr(ops, k), k
        setattr(ops, k, type(getattr(ops, k))(int(v)))
    ms = timed()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    if first is None:
        first = grads
    worst = max((float((g - first[k]).abs().max()) / max(float(first[k].abs().max()), 2e-3), k) for k, g in grads.items())
    print('%-40s %.3f ms fwd+bwd   worst gradient delta vs first arm %.2e (%s)' % (arm, ms, worst[0], worst[1]), flush=True)
