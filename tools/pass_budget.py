"""Pass-budget experiment: for every large tensor-core contraction site of the step, run the
full-size training step (B = 64 x 1800, config #3) with that site on ONE TF32 pass (precision = 2) and report, against
the all-3xTF32 run on the same inputs: logits error, arg-max(S) flips, worst parameter-gradient error (relative to
each tensor's largest entry) and the step time.  Usage (GPU box): python tools/pass_budget.py [--norm-adj]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cgc_net_b200 import graph, network, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--norm-adj', action='store_true')
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'pass_budget.json'))
args = ap.parse_args()

dev = torch.device('cuda:0')
graphs = graph.synthetic_batch(args.batch, 1800, 8, 16, seed=5)
batch = graph.Batch.from_data_list(graphs).to(dev)
torch.manual_seed(1)
model = network.SoftPoolingGcnEncoder(11404, 16, 20, 20, True, True, 20, 3, 0.1, [50], concat=True,
                                      load_data_sparse=True, collect_assign=True, norm_adj=args.norm_adj).to(dev).train()
state = {k: v.clone() for k, v in model.state_dict().items()}


def run():
    model.load_state_dict(state)
    model.zero_grad(set_to_none=True)
    logits, loss = model(batch)
    loss.backward()
    torch.cuda.synchronize()
    return (logits.detach().clone(), [a.clone() for a in model.assign_index],
            {k: p.grad.clone() for k, p in model.named_parameters()})


def timed(reps=3):
    run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model.zero_grad(set_to_none=True)
        _, loss = model(batch)
        loss.backward()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


sites = sorted(ops.PASS_BUDGET)
shipped = dict(ops.PASS_BUDGET)
for s in sites:
    ops.PASS_BUDGET[s] = None
ref_logits, ref_arg, ref_grads = run()
t_ref = timed()
rows = [{'site': 'all 3xTF32', 'ms': t_ref}]
print('all 3xTF32: %.2f ms fwd+bwd' % t_ref)


def report(name):
    logits, arg, grads = run()
    ms = timed()
    lerr = float((logits - ref_logits).abs().max() / ref_logits.abs().max())
    flips = sum(int((a != b).sum()) for a, b in zip(arg, ref_arg))
    worst, wname = 0.0, ''
    for k, gr in grads.items():
        r = ref_grads[k]
        e = float((gr - r).abs().max()) / max(float(r.abs().max()), 2e-3)
        if e > worst:
            worst, wname = e, k
    rows.append({'site': name, 'ms': ms, 'logits_rel': lerr, 'argmax_flips': flips, 'worst_grad': worst,
                 'worst_grad_tensor': wname})
    print('%-28s %.2f ms  logits %.2e  argmax flips %d  worst grad %.2e (%s)' % (name, ms, lerr, flips, worst, wname))


for s in sites:
    ops.PASS_BUDGET[s] = 2
    report(s)
    ops.PASS_BUDGET[s] = None
bwd = [s for s in sites if s in ('lin_dw', 'lin_dx', 'pool_gs', 'pool_gp', 'pool_gx', 'pool_gadj', 'dense_agg_t',
                                 'dense_dadj')]
for s in bwd:
    ops.PASS_BUDGET[s] = 2
report('all backward sites')
for s in ('pool_stas', 'pool_stx'):
    ops.PASS_BUDGET[s] = 2
report('backward + pool forward')
for s in sites:
    ops.PASS_BUDGET[s] = shipped[s]
report('shipped budget')
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(rows, open(args.out, 'w'), indent=1)
