"""Benchmark of the CGC-Net graph-convolution hot path on B200 (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one training pass of SoftPoolingGcnEncoder over one batch of synthetic cell graphs:
CSR build from the COO edge list, forward, backward, gradient all-reduce (N > 1) and the Adam update
(reference train.py:178-183).  Workload = BASELINE.json configs[2]: 64 graphs of 1800 nodes, degree 8,
16 features per GPU (weak scaling), model SoftPoolingGcnEncoder(11404, 16, 20, 20, ..., 0.10) ->
1140 / 114 clusters, fp32.

One JSON line is printed by rank 0:
  value      graphs/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e        the same through the public API with HOST (pinned) inputs: per-step H2D copy of x / edge_index /
             batch / y (graph.DevicePrefetcher: batch i+1 is copied on a side stream while batch i computes) and
             a D2H read of the loss, all inside the timed region
  roofline   the SAGE gather kernel (cgc_sage_gather_ell_fwd) alone: algorithmic bytes / launch time
             vs the measured HBM copy bandwidth (MEASURED_PEAKS.json); `traffic` = DRAM bytes of one launch
             from the committed ncu capture (profiles/gather_r1_ncu.txt)
  roofline_tensor  the kernel with the largest share of the step (tcgen05 3xTF32 GEMM, CTA pairs): algorithmic
             fp32 FLOPs / launch time vs half the measured bf16 peak, and the same counting the 3 MMA passes
  cpu_baseline  the dense CPU oracle port of the reference path timed on this box's host cores on a
             bounded sample (N = 1, rank 0 only)
`--impl reference` times that CPU oracle port alone (the reference is pure Python on dense torch ops;
torch_geometric is not installable offline) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL_ARGS = (11404, 16, 20, 20, True, True, 20, 3, 0.10, [50])      # train.py:254-261, setting.py:15
METRIC = 'graphs/sec (1800-node, deg-8) fwd+bwd'
FEAT = 16
# dram__bytes_read.sum + dram__bytes_write.sum of one k_sage_gather_ell launch at B=64 x 1800 x F=16, cold L2
# (ncu --set full, profiles/gather_r1_ncu.txt): 9.77 MB read (tile 7.37 + uint16 indices 1.84 + scales / map),
# 0 written inside the launch (the 7.37 MB of outputs are still in L2 when it ends).
GATHER_NCU_DRAM_BYTES = 9768704


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch-per-gpu', type=int, default=64)
    ap.add_argument('--nodes', type=int, default=1800)
    ap.add_argument('--degree', type=int, default=8)
    ap.add_argument('--norm-adj', action='store_true')
    ap.add_argument('--cpu-sample-graphs', type=int, default=16)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    return ap.parse_args()


def workload_name(args):
    return 'batch %d x %d-node deg-%d graphs, %d features, full SoftPoolingGcnEncoder fwd+bwd+Adam%s' % (
        args.batch_per_gpu, args.nodes, args.degree, FEAT, ' (norm_adj)' if args.norm_adj else '')


# ----------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_rate(args, steps, warmup):
    """graphs/s of the dense CPU oracle port (fwd + bwd + Adam) on a bounded sample of the workload."""
    from cgc_net_b200 import graph
    from oracle import restatement
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_graphs = args.cpu_sample_graphs
    graphs = graph.synthetic_batch(n_graphs, args.nodes, args.degree, FEAT, seed=0)
    batch = graph.Batch.from_data_list(graphs)
    torch.manual_seed(0)
    model = restatement.OracleEncoder(*MODEL_ARGS, concat=True, gcn_name='SAGE', load_data_sparse=True,
                                      norm_adj=args.norm_adj).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)

    def one():
        opt.zero_grad()
        _, loss = model(batch)
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return n_graphs / dt, dt, cores, n_graphs


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))      # ~10 s of CPU work on 16 cores
    rate, dt, cores, n_graphs = cpu_oracle_rate(args, steps, warmup)
    sample = '%d of the %d graphs per step (dense CPU oracle port, fwd+bwd+Adam), %d steps' % (
        n_graphs, args.batch_per_gpu, steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'graphs/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'timed_sample': sample},
        'cpu_baseline': {'value': rate, 'unit': 'graphs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'graphs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- gather roofline
def gather_roofline(args, device, variant='ell', parity=True):
    """The SAGE gather kernel alone (F = 16, unfused): CUDA-graph replay over rotating input copies whose
    total footprint exceeds L2, timed with CUDA events on the launching stream.
    variant 'ell' = cgc_sage_gather_ell_fwd (fixed-pitch index table), 'csr' = cgc_sage_gather_mean_fwd."""
    from cgc_net_b200 import _backend, graph, ops
    be = _backend.get()
    b, n, k = args.batch_per_gpu, args.nodes, args.degree
    graphs = graph.synthetic_batch(min(b, 8), n, k, FEAT, seed=1)
    graphs = (graphs * ((b + len(graphs) - 1) // len(graphs)))[:b]
    batch = graph.Batch.from_data_list(graphs).to(device)
    xp, g = ops.build_sparse_graph(batch.x, batch.edge_index, batch.batch, batch.num_nodes_list, False)
    nnz = int(g.rowptr[-1])
    rows = g.rows
    alg_bytes = 4 * rows * FEAT + 4 * nnz + 4 * (rows + b) + 4 * rows * FEAT     # SURVEY 8(d) per-graph formula x B
    per_copy = alg_bytes
    copies = max(2, int(320e6 // per_copy) + 1)                                # > 2x the 126 MB L2
    xs = [xp.clone() for _ in range(copies)]
    outs = [torch.empty_like(xp) for _ in range(copies)]
    if variant == 'ell':
        assert g.ell is not None, 'no ELL16 table was built for this batch'
        ell0 = g.ell
        if not parity:
            ell0, perm0, sc0 = torch.empty_like(g.ell), torch.empty_like(g.ell_perm), torch.empty_like(g.ell_scale)
            be.csr_to_ell16(g.rowptr, g.col, None, g.rs, b, n, g.ell_width, False, ell0, None, sc0, perm0)
        else:
            perm0, sc0 = g.ell_perm, g.ell_scale
        ells = [ell0.clone() for _ in range(copies)]
        perms = [perm0.clone() for _ in range(copies)]
        scs = [sc0.clone() for _ in range(copies)]

        def launch(i):
            be.sage_gather_ell_fwd(ells[i], None, scs[i], perms[i], g.ell_width, g.rowptr, g.col, None, xs[i], b, n,
                                   outs[i])
    else:
        cols = [g.col.clone() for _ in range(copies)]
        ptrs = [g.rowptr.clone() for _ in range(copies)]

        def launch(i):
            be.sage_gather_mean_fwd(ptrs[i], cols[i], None, None, xs[i], b, n, outs[i])

    launch(0)
    torch.cuda.synchronize()
    if not os.environ.get('CGC_GATHER_STAGE'):
        ref = torch.empty_like(xp)
        be.spmm(g.rowptr, g.col, None, xp, None, g.rs, ref, False)
        assert torch.allclose(outs[0], ref, rtol=1e-5, atol=1e-6), 'gather kernel disagrees with the generic gather'
    stream = torch.cuda.Stream(device)
    stream.wait_stream(torch.cuda.current_stream())
    graph_obj = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for i in range(copies):
            launch(i)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph_obj, stream=stream):
            for i in range(copies):
                launch(i)
    reps, blocks = 20, 5
    samples = []
    with torch.cuda.stream(stream):
        for _ in range(3):
            graph_obj.replay()
        for _ in range(blocks):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(reps):
                graph_obj.replay()
            e1.record(stream)
            torch.cuda.synchronize()
            samples.append(e0.elapsed_time(e1) * 1e3 / (reps * copies))
    samples.sort()
    gather_roofline.samples = samples
    return alg_bytes, samples[len(samples) // 2], copies       # median of 5 blocks of 20 x `copies` launches


def gemm_roofline(device):
    """The kernel with the largest share of the step (k_gemm_tc, 53 %): the `lin` contraction of GCN_pool_1
    (115200 x 1140 x 1180, network.py:121-124) alone, CUDA events around 5 launches after 3 warm-ups; the three
    operands (1.6 GB) exceed L2.  Algorithmic FLOPs = 2 M N K (what the reference's fp32 matmul computes); the
    kernel spends 3 tf32 MMA passes on them (3xTF32)."""
    from cgc_net_b200 import _backend
    be = _backend.get()
    m, n, k = 115200, 1140, 1180
    a = torch.randn(m, k, device=device)
    b = torch.randn(n, k, device=device)
    c = torch.empty(m, n, device=device)
    for _ in range(3):
        be.gemm(a, b, c, False, True, precision=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        be.gemm(a, b, c, False, True, precision=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return 2.0 * m * n * k, ms


# ----------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from cgc_net_b200 import _backend, graph, network, parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback for the product path)'
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    be = _backend.get()

    b = args.batch_per_gpu
    graphs = graph.synthetic_batch(b, args.nodes, args.degree, FEAT, seed=100 + rank)
    host_batch = graph.Batch.from_data_list(graphs).pin_memory()
    dev_batch = host_batch.to(device)
    torch.manual_seed(0)
    model = network.SoftPoolingGcnEncoder(*MODEL_ARGS, concat=True, gcn_name='SAGE', load_data_sparse=True,
                                          norm_adj=args.norm_adj).to(device).train()
    parallel.broadcast_parameters(model)
    bucket = parallel.FlatGradBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)

    def step(batch):
        bucket.zero()
        _, loss = model(batch)
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(3, args.warmup)):
        step(dev_batch)
    sampler = ClockSampler(local) if rank == 0 else None
    k0 = be.kernel_launches()
    ms_dev = timed(lambda: step(dev_batch), args.steps)
    launches = be.kernel_launches() - k0

    # end to end through the public API: host (pinned) batches go through graph.DevicePrefetcher (copy of batch i+1
    # on a side stream while batch i computes), every step's H2D copy and the D2H read of its loss are inside
    # the timed region
    def e2e_run(steps):
        # the loss of every step is copied device->host (pinned) as part of that step, without stalling the launch
        # queue; the host looks at the values once they have all landed (end of the timed region)
        losses = torch.empty(steps, dtype=torch.float32).pin_memory()
        for i, batch in enumerate(graph.DevicePrefetcher([host_batch] * steps, device)):
            losses[i:i + 1].copy_(step(batch).detach().view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        assert bool(torch.isfinite(losses).all()), 'non-finite loss in the end-to-end run'
        return losses

    e2e_run(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_t) / args.steps
    clocks = sampler.stop() if sampler is not None else None

    total_graphs = b * world
    line = {
        'metric': METRIC, 'value': total_graphs / (ms_dev * 1e-3), 'unit': 'graphs/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_dev, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'graphs_per_gpu': b, 'global_graphs': total_graphs,
                   'clusters': [1140, 114], 'parallelism': 'dp%d (batch-sharded, one flat-bucket all-reduce)' % world,
                   'l2': 'working set >> L2: ~%d MB of [M,1140] fp32 activations per tensor, > 10 GB per step'
                         % (b * args.nodes * 1140 * 4 // 2 ** 20)},
        'e2e': {'value': total_graphs / (ms_e2e * 1e-3), 'unit': 'graphs/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': host_batch.nbytes(), 'd2h_bytes_per_step': 4},
        'gpu_launches': launches,
        'clocks': clocks,
        'grad_allreduce_bytes': bucket.nbytes() if world > 1 else 0,
    }
    if rank == 0 and not args.no_roofline:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.isfile(peaks_path):
            peak, which = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            peak, which = 6650.0, 'fallback (B200_PROFILING.md)'
        alg_bytes, us, copies = gather_roofline(args, device)
        achieved = alg_bytes / (us * 1e-6) / 1e9
        line['roofline'] = {'kernel': 'k_sage_gather_ell (cgc_sage_gather_ell_fwd, F=16, unfused)', 'bound': 'hbm',
                            'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                            'traffic': GATHER_NCU_DRAM_BYTES if (args.batch_per_gpu, args.nodes, args.degree) == (64, 1800, 8) else None,
                            'traffic_source': 'profiles/gather_r1_ncu.txt (ncu --set full, one cold launch; '
                                              'outputs, 7.37 MB, still in L2 at kernel end)',
                            'peak_source': which, 'algorithmic_bytes_per_launch': alg_bytes,
                            'us_per_launch': us,
                            'timing': 'CUDA-graph replay of %d launches over %d rotating input copies (%.0f MB > L2), '
                                      'CUDA events on the launching stream, median of 5 blocks of 20 replays'
                                      % (copies, copies, copies * alg_bytes / 1e6)}
    if rank == 0 and not args.no_roofline:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.isfile(peaks_path):
            tpeak, twhich = float(json.load(open(peaks_path))['bf16_tflops']) / 2.0, \
                'measured (MEASURED_PEAKS.json bf16_tflops / 2: tf32 runs at half the bf16 rate)'
        else:
            tpeak, twhich = 1590.0 / 2.0, 'fallback (B200_PROFILING.md 1.59 PFLOP/s bf16 / 2)'
        flops, ms = gemm_roofline(device)
        line['roofline_tensor'] = {
            'kernel': 'k_gemm_tc<128, pair> (cgc_gemm, tcgen05 3xTF32, cta_group::2) on lin 115200x1140x1180',
            'bound': 'tensor', 'achieved': flops / ms / 1e9, 'peak': tpeak, 'unit': 'TFLOP/s',
            'frac': flops / ms / 1e9 / tpeak, 'mma_passes': 3, 'mma_work_frac': 3.0 * flops / ms / 1e9 / tpeak,
            'peak_source': twhich, 'ms_per_launch': ms, 'step_share': 0.53,
            'note': 'achieved = algorithmic fp32 FLOPs (2MNK) per second; the kernel issues 3 tf32 MMA passes per '
                    'product to stay within fp32 parity, mma_work_frac counts those'}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, dt, cores, n_graphs = cpu_oracle_rate(args, 8, 1)                     # ~8 s of CPU work on 16 cores
        line['cpu_baseline'] = {'value': rate, 'unit': 'graphs/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d of the %d graphs per step (dense CPU oracle port of the reference '
                                          'path, fwd+bwd+Adam), 8 timed steps after 1 warm-up' % (n_graphs, b)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
