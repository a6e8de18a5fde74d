"""One profiled training step of the bench workload (for ncu: --profile-from-start off)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cgc_net_b200 import graph, network, parallel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--norm-adj', action='store_true')
a = ap.parse_args()
dev = torch.device('cuda', 0)
graphs = graph.synthetic_batch(min(a.batch, 8), 1800, 8, 16, seed=0)
graphs = (graphs * ((a.batch + len(graphs) - 1) // len(graphs)))[:a.batch]
batch = graph.Batch.from_data_list(graphs).to(dev)
torch.manual_seed(0)
model = network.SoftPoolingGcnEncoder(*bench.MODEL_ARGS, concat=True, load_data_sparse=True,
                                      norm_adj=a.norm_adj).to(dev).train()
bucket = parallel.FlatGradBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)


def step():
    bucket.zero()
    _, loss = model(batch)
    loss.backward()
    opt.step()


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('profiled', a.steps, 'step(s)')
