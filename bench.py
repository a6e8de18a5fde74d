"""Benchmark of the CGC-Net graph-convolution hot path on B200 (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3|4|5]

A "step" is one training pass of SoftPoolingGcnEncoder over one batch of synthetic cell graphs:
CSR build from the COO edge list, forward, backward, gradient all-reduce (N > 1) and the Adam update
(reference train.py:178-183).  Model SoftPoolingGcnEncoder(11404, 16, 20, 20, ..., 0.10) -> 1140 / 114 clusters,
fp32.  Workloads (BASELINE.json configs):
  --config 3 (default)  64 graphs of 1800 nodes, degree 8, 16 features PER GPU (weak scaling) -- the headline
  --config 4            256 graphs of 1800 nodes in total, sharded 256 / N per GPU (strong scaling)
  --config 5            32 ragged graphs per GPU (weak scaling): 1000-6000 raw nuclei, sub-sampled to 500-3000
                        nodes with the reference's `fuse` / `random` samplers (alternating), edges from
                        radius_graph(pos, 100, loop=True, k) with k ~ U{5..12} per graph, node-balanced split
                        across ranks (PyG DataParallel.scatter) -- all of it on the GPU (graph.sample_nodes,
                        graph.radius_graph), timed separately as `prep_ms`

One JSON line is printed by rank 0:
  value      graphs/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e        the same through the public API with HOST (pinned) inputs: per-step H2D copy of x / edge_index /
             batch / y (graph.DevicePrefetcher: batch i+1 is copied on a side stream while batch i computes) and
             a D2H read of the loss, all inside the timed region
  roofline   the SAGE gather kernel (cgc_sage_gather_ell, F=16, unfused -- the BASELINE metric) alone: algorithmic
             bytes / launch time vs the measured HBM copy bandwidth (MEASURED_PEAKS.json); `traffic` = DRAM bytes of
             one launch read from profiles/gather_traffic.json (written by tools/ncu_summary.py from an ncu capture);
             `variants` = the same kernel on a radius_graph-style (variable row degree) adjacency and on the
             weighted (norm_adj) table, and the fused layer kernel (cgc_sage_layer_ell_fwd, 16 -> 20)
  roofline_tensor  the kernel with the largest share of the step (tcgen05 GEMM, CTA pairs) on the `lin` contraction
  cpu_baseline  the reference's CPU path timed on this box's host cores on a bounded sample (N = 1, rank 0 only):
             the VERBATIM reference (kind "reference") when a checkout is reachable (/root/reference or
             baseline/_ref), else the dense CPU oracle port (kind "port")
`--impl reference` times that CPU path alone and prints the same line shape.
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
TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'gather_traffic.json')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=3, choices=[3, 4, 5])
    ap.add_argument('--batch-per-gpu', type=int, default=None)
    ap.add_argument('--nodes', type=int, default=1800)
    ap.add_argument('--degree', type=int, default=8)
    ap.add_argument('--norm-adj', action='store_true')
    ap.add_argument('--cpu-sample-graphs', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--overlap-allreduce', type=int, default=0, metavar='CHUNKS',
                    help='N > 1 ranks: all-reduce the gradient bucket in CHUNKS pieces started from inside backward() '
                         '(parallel.FlatGradBucket.enable_overlap; gloo-tested, off by default: not yet measured on NCCL)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.batch_per_gpu is None:
        args.batch_per_gpu = {3: 64, 4: max(1, 256 // world), 5: 32}[args.config]
    return args


def workload_name(args):
    tail = ', 16 features, full SoftPoolingGcnEncoder fwd+bwd+Adam%s' % (' (norm_adj)' if args.norm_adj else '')
    if args.config == 4:
        return 'config #4: 256 x %d-node deg-%d graphs in total, sharded by batch index%s' % (args.nodes, args.degree, tail)
    if args.config == 5:
        return ('config #5: %d ragged graphs per GPU, 500-3000 nodes after fuse/random sub-sampling (ratio 0.5), '
                'radius_graph degree cap 5-12, node-balanced split%s' % (args.batch_per_gpu, tail))
    return 'batch %d x %d-node deg-%d graphs%s' % (args.batch_per_gpu, args.nodes, args.degree, tail)


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


# ----------------------------------------------------------------------------------- workloads
def host_graphs(args, rank, world):
    """This rank's list of graph.Data (host tensors) for --config 3 / 4."""
    from cgc_net_b200 import graph
    if args.config == 4:
        per = args.batch_per_gpu
        # the SAME 256 graphs whatever N (strong scaling): graph i has seed 1000 + i; rank r owns [r*per, (r+1)*per)
        return [graph.synthetic_cell_graph(args.nodes, args.degree, FEAT, seed=4000003 + rank * per + i)
                for i in range(per)]
    return graph.synthetic_batch(args.batch_per_gpu, args.nodes, args.degree, FEAT, seed=100 + rank)


def ragged_graphs_on_gpu(args, rank, world, device):
    """--config 5: raw point sets -> fuse / random node sub-sampling -> radius_graph, on the GPU (the loader-side
    rows f-2 of SURVEY 8), then PyG's node-balanced contiguous split; returns (this rank's Data list, prep ms)."""
    from cgc_net_b200 import graph
    total = args.batch_per_gpu * world
    gen = torch.Generator().manual_seed(555)
    raws, ks = [], []
    for i in range(total):
        n_raw = int(torch.randint(1000, 6001, (1,), generator=gen))
        ks.append(int(torch.randint(5, 13, (1,), generator=gen)))
        raws.append(graph.synthetic_cell_graph(n_raw, 1, FEAT, seed=900000 + i))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = []
    for method in ('fuse', 'random'):
        sel = [i for i in range(total) if (i % 2 == 0) == (method == 'fuse')]
        if not sel:
            continue
        counts = [raws[i].num_nodes for i in sel]
        pos = torch.cat([raws[i].pos for i in sel]).to(device)
        ids = graph.sample_nodes(pos, counts, 0.5, method, generator=gen)
        off = 0
        for i, n, keep in zip(sel, counts, ids):
            keep = keep.sort().values
            p = pos[off:off + n][keep]
            ei = graph.radius_graph(p, 100.0, None, True, ks[i], num_nodes_list=[int(keep.numel())])
            out.append((i, graph.Data(x=raws[i].x[keep.cpu()], edge_index=ei.cpu(), y=raws[i].y, pos=p.cpu())))
            off += n
    torch.cuda.synchronize()
    prep_ms = (time.perf_counter() - t0) * 1e3
    data = [d for _, d in sorted(out, key=lambda t: t[0])]
    shard = graph.split_by_node_count(data, world)[rank]
    return shard, prep_ms


# ----------------------------------------------------------------------------------- CPU reference arm
def reference_root():
    for cand in (os.environ.get('CGC_REFERENCE_ROOT'), '/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if cand and os.path.isfile(os.path.join(cand, 'model', 'network.py')):
            return cand
    return None


def cpu_reference_rate(args, steps, warmup, n_graphs):
    """graphs/s of the reference's CPU path (fwd + bwd + Adam, all host threads) on `n_graphs` graphs of the
    workload per step.  Verbatim reference when a checkout is reachable, else the oracle port."""
    from cgc_net_b200 import graph
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nodes, degree = (args.nodes, args.degree)
    if args.config == 5:
        graphs = graph.synthetic_batch(n_graphs, ragged=(500, 3000, 5, 12), feat_dim=FEAT, seed=0)
    else:
        graphs = graph.synthetic_batch(n_graphs, nodes, degree, FEAT, seed=0)
    batch = graph.Batch.from_data_list(graphs)
    torch.manual_seed(0)
    root = reference_root()
    kw = dict(concat=True, gcn_name='SAGE', load_data_sparse=True, norm_adj=args.norm_adj)
    if root is not None:
        os.environ['CGC_REFERENCE_ROOT'] = root
        from oracle import reference_loader
        reference_loader.REFERENCE_ROOT = root
        model = reference_loader.build_reference_encoder(*MODEL_ARGS, cpu=True, **kw).train()
        kind = 'reference'
        forward = lambda: reference_loader.run_reference(model, batch)      # noqa: E731
    else:
        from oracle import restatement
        model = restatement.OracleEncoder(*MODEL_ARGS, **kw).train()
        kind = 'port'
        forward = lambda: model(batch)                                      # noqa: E731
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)

    def one():
        opt.zero_grad()
        _, loss = forward()
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return n_graphs / dt, dt, cores, kind


def cpu_sample_size(args, steps, warmup, budget_s):
    """Graphs per CPU step so that (steps + warmup) steps take about `budget_s` seconds at ~18 graphs/s
    (16 host cores); never more than the real per-GPU batch."""
    if args.cpu_sample_graphs:
        return min(args.cpu_sample_graphs, args.batch_per_gpu)
    fit = int(budget_s * 18.0 / max(1, steps + warmup))
    return max(4, min(args.batch_per_gpu, 64, fit))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    n_graphs = cpu_sample_size(args, steps, warmup, 200.0)
    rate, dt, cores, kind = cpu_reference_rate(args, steps, warmup, n_graphs)
    what = 'verbatim reference model/network.py behind the PyG shim' if kind == 'reference' else 'dense CPU oracle port'
    sample = '%d of the %d graphs per step (%s, fwd+bwd+Adam, %d host threads), %d timed steps after %d warm-ups' % (
        n_graphs, args.batch_per_gpu, what, cores, steps, warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'graphs/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'strong' if args.config == 4 else 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'timed_sample': sample},
        'cpu_baseline': {'value': rate, 'unit': 'graphs/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': rate, 'unit': 'graphs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- gather roofline
def gather_roofline(args, device, kind='gather', row_is_centre=True, norm_adj=False, pdl=True, check=True,
                    two_per_sm=None, multicast=True):
    """One kernel of the gather family alone at the BASELINE shape (B x nodes x F = 16): CUDA-graph replay over
    rotating input copies whose total footprint exceeds L2, timed with CUDA events on the launching stream.

    kind 'gather' = cgc_sage_gather_ell (unfused, the BASELINE metric); 'layer' = cgc_sage_layer_ell_fwd (16 -> 20,
    gather + W + bias + L2 norm + mask + ReLU + BatchNorm partial sums + finalize); 'csr' = the round-1a CSR kernel.
    row_is_centre False = radius_graph-style adjacency (fixed column degree, variable row degree: pitch-16 table,
    rows above the pitch walk the CSR); norm_adj = weighted table (_re_norm_adj values).
    Returns (algorithmic bytes per launch, us per launch, copies)."""
    from cgc_net_b200 import _backend, graph, ops
    be = _backend.get()
    b, n, k = args.batch_per_gpu, args.nodes, args.degree
    graphs = graph.synthetic_batch(min(b, 8), n, k, FEAT, seed=1, row_is_centre=row_is_centre)
    graphs = (graphs * ((b + len(graphs) - 1) // len(graphs)))[:b]
    batch = graph.Batch.from_data_list(graphs).to(device)
    xp, g = ops.build_sparse_graph(batch.x, batch.edge_index, batch.batch, batch.num_nodes_list, norm_adj)
    nnz = int(g.rowptr[-1])
    rows = g.rows
    fout = 20 if kind == 'layer' else FEAT
    # SURVEY 8(d) per-graph formula x B: X once, int32 column indices, rowptr, [values], output [+ W, b]
    alg_bytes = 4 * rows * FEAT + 4 * nnz + 4 * (rows + b) + (4 * nnz if norm_adj else 0) + 4 * rows * fout
    if kind == 'layer':
        alg_bytes += b * 1360
    copies = max(2, int(320e6 // alg_bytes) + 1)                                # > 2x the 126 MB L2
    xs = [xp.clone() for _ in range(copies)]
    outs = [torch.empty(rows, FEAT, device=device) for _ in range(copies)]
    if kind in ('gather', 'layer'):
        assert g.ell is not None, 'no ELL16 table was built for this batch'
        ells = [g.ell.clone() for _ in range(copies)]
        vals = [g.ell_val.clone() if g.ell_val is not None else None for _ in range(copies)]
        perms = [g.ell_perm.clone() for _ in range(copies)]
        scs = [g.ell_scale.clone() for _ in range(copies)]
    if kind == 'gather':
        def launch(i):
            # binary adjacency: no scale vector (mean derived from the row's slots), as ops._agg calls it
            be.sage_gather_ell(ells[i], vals[i], scs[i] if vals[i] is not None else None, perms[i], g.ell_width,
                               g.rowptr, g.col, g.val, xs[i], b, n, outs[i], tables_ready=True, pdl=pdl,
                               two_per_sm=two_per_sm, multicast=multicast)
    elif kind == 'layer':
        torch.manual_seed(3)
        w = (torch.rand(FEAT, 20, device=device) - 0.5) * 0.5
        bias = (torch.rand(20, device=device) - 0.5) * 0.5
        ys = [torch.empty(rows, 20, device=device) for _ in range(copies)]
        rm, rv = torch.zeros(20, device=device), torch.ones(20, device=device)
        mean, invstd = torch.empty(20, device=device), torch.empty(20, device=device)

        def launch(i):
            be.sage_layer_ell_fwd(ells[i], vals[i], scs[i] if vals[i] is not None else None, perms[i], g.ell_width,
                                  g.rowptr, g.col, g.val, xs[i], b, n,
                                  w, bias, g.mask, 0, outs[i], ys[i], rm, rv, 0.1, 1e-5, mean, invstd,
                                  tables_ready=True, pdl=pdl)
    else:
        cols = [g.col.clone() for _ in range(copies)]
        ptrs = [g.rowptr.clone() for _ in range(copies)]

        def launch(i):
            be.sage_gather_mean_fwd(ptrs[i], cols[i], g.val, None if g.val is None else g.rs, xs[i], b, n, outs[i])

    launch(0)
    torch.cuda.synchronize()
    if check:
        ref = torch.empty_like(xp)
        be.spmm(g.rowptr, g.col, g.val, xp, None, g.rs, ref, False)
        assert torch.allclose(outs[0], ref, rtol=1e-5, atol=1e-6), 'gather kernel disagrees with the generic gather'
        if kind == 'layer':
            yref = torch.empty(rows, 20, device=device)
            be.sage_epilogue_fwd(ref, w, bias, g.mask, yref, 0)
            assert torch.allclose(ys[0], yref, rtol=1e-4, atol=1e-6), 'fused layer disagrees with gather + epilogue'
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


def gemm_roofline(device, precision=1):
    """The kernel with the largest share of the step (k_gemm_tc): the `lin` contraction of GCN_pool_1
    (115200 x 1140 x 1180, network.py:121-124) alone, CUDA events around 5 launches after 3 warm-ups; the three
    operands (1.6 GB) exceed L2.  Algorithmic FLOPs = 2 M N K (what the reference's fp32 matmul computes);
    precision 1 spends 3 tf32 MMA passes on them (3xTF32), precision 2 one."""
    from cgc_net_b200 import _backend
    be = _backend.get()
    m, n, k = 115200, 1140, 1180
    a = torch.randn(m, k, device=device)
    b = torch.randn(n, k, device=device)
    c = torch.empty(m, n, device=device)
    for _ in range(3):
        be.gemm(a, b, c, False, True, precision=precision)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        be.gemm(a, b, c, False, True, precision=precision)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return 2.0 * m * n * k, ms


def roofline_lines(args, device):
    """`roofline` (the BASELINE gather) with its variants, and `roofline_tensor`."""
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(peaks_path):
        pk = json.load(open(peaks_path))
        peak, which = float(pk['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        tpeak, twhich = float(pk['bf16_tflops']) / 2.0, \
            'measured (MEASURED_PEAKS.json bf16_tflops / 2: tf32 runs at half the bf16 rate)'
    else:
        peak, which = 6650.0, 'fallback (B200_PROFILING.md)'
        tpeak, twhich = 1590.0 / 2.0, 'fallback (B200_PROFILING.md 1.59 PFLOP/s bf16 / 2)'
    rargs = argparse.Namespace(batch_per_gpu=64, nodes=1800, degree=8)       # always the BASELINE shape
    out = {}
    alg, us, copies = gather_roofline(rargs, device)
    achieved = alg / (us * 1e-6) / 1e9
    traffic, tsrc = None, 'no ncu capture of this kernel build committed'
    if os.path.isfile(TRAFFIC_FILE):
        t = json.load(open(TRAFFIC_FILE))
        traffic, tsrc = t.get('dram_bytes'), t.get('source')
    variants = {}
    for name, kw in (('unfused, plain stream-ordered launch (no PDL)', dict(pdl=False)),
                     ('unfused, one 1024-thread CTA per SM (k_sage_gather_ell; no early L2 prefetch of the tile)',
                      dict(two_per_sm=False)),
                     ('unfused, two CTAs per SM, every CTA copying its graph\'s whole tile itself (no cluster multicast)',
                      dict(multicast=False)),
                     ('unfused, radius_graph-style adjacency (variable row degree, pitch-16 table)',
                      dict(row_is_centre=False)),
                     ('unfused, weighted table (norm_adj)', dict(norm_adj=True)),
                     ('fused layer 16->20 (gather + W + bias + L2 norm + mask + ReLU + BN sums)', dict(kind='layer')),
                     ('fused layer 16->20, weighted table (norm_adj)', dict(kind='layer', norm_adj=True))):
        try:
            a2, u2, _ = gather_roofline(rargs, device, **kw)
            variants[name] = {'us_per_launch': u2, 'algorithmic_bytes_per_launch': a2,
                              'achieved': a2 / (u2 * 1e-6) / 1e9, 'frac': a2 / (u2 * 1e-6) / 1e9 / peak}
        except Exception as exc:                                         # a variant must not take the line down
            variants[name] = {'error': str(exc)[:200]}
    out['roofline'] = {
        'kernel': 'k_sage_gather_ell_lean<2> (cgc_sage_gather_ell, F=16, unfused, fixed degree 8; two 512-thread CTAs '
                  'per SM, the two CTAs of a graph are a cluster sharing one multicast copy of the tile, the next '
                  'launch pulls its tile into L2 while this one gathers)', 'bound': 'hbm',
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
        'traffic_source': tsrc, 'peak_source': which, 'algorithmic_bytes_per_launch': alg, 'us_per_launch': us,
        'timing': 'CUDA-graph replay of %d launches over %d rotating input copies (%.0f MB > L2), programmatic '
                  'dependent launch between consecutive launches (as in the training step), CUDA events on the '
                  'launching stream, median of 5 blocks of 20 replays' % (copies, copies, copies * alg / 1e6),
        'variants': variants}
    flops, ms = gemm_roofline(device, 1)
    flops1, ms1 = gemm_roofline(device, 2)
    out['roofline_tensor'] = {
        'kernel': 'k_gemm_tc<128, pair> (cgc_gemm, tcgen05 3xTF32, cta_group::2) on lin 115200x1140x1180',
        'bound': 'tensor', 'achieved': flops / ms / 1e9, 'peak': tpeak, 'unit': 'TFLOP/s',
        'frac': flops / ms / 1e9 / tpeak, 'mma_passes': 3, 'mma_work_frac': 3.0 * flops / ms / 1e9 / tpeak,
        'peak_source': twhich, 'ms_per_launch': ms,
        'single_pass': {'ms_per_launch': ms1, 'achieved': flops1 / ms1 / 1e9, 'frac': flops1 / ms1 / 1e9 / tpeak},
        'note': 'achieved = algorithmic fp32 FLOPs (2MNK) per second; 3xTF32 issues 3 tf32 MMA passes per product to '
                'stay within fp32 parity (mma_work_frac counts those); single_pass = precision 2 (pass budget)'}
    return out


# ----------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from cgc_net_b200 import _backend, graph, network, ops, parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback for the product path)'
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    be = _backend.get()

    prep_ms = None
    if args.config == 5:
        graphs, prep_ms = ragged_graphs_on_gpu(args, rank, world, device)
    else:
        graphs = host_graphs(args, rank, world)
    b = len(graphs)
    host_batch = graph.Batch.from_data_list(graphs).pin_memory()
    dev_batch = host_batch.to(device)
    torch.manual_seed(0)
    model = network.SoftPoolingGcnEncoder(*MODEL_ARGS, concat=True, gcn_name='SAGE', load_data_sparse=True,
                                          norm_adj=args.norm_adj).to(device).train()
    parallel.broadcast_parameters(model)
    bucket = parallel.FlatGradBucket(model.parameters())
    overlapped = bool(args.overlap_allreduce > 0 and world > 1 and bucket.enable_overlap(args.overlap_allreduce))
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

    counts = torch.tensor([b, host_batch.x.shape[0], host_batch.nbytes()], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(counts)
    total_graphs, total_nodes, total_h2d = int(counts[0]), int(counts[1]), int(counts[2])
    nmax = max(host_batch.num_nodes_list)
    budget = {k: v for k, v in ops.PASS_BUDGET.items() if v is not None}
    line = {
        'metric': METRIC, 'value': total_graphs / (ms_dev * 1e-3), 'unit': 'graphs/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_dev, 'higher_is_better': True,
        'scaling': 'strong' if args.config == 4 else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'bench_config': args.config, 'graphs_rank0': b,
                   'global_graphs': total_graphs, 'global_nodes': total_nodes, 'clusters': [1140, 114],
                   'parallelism': 'dp%d (batch-sharded, %s)' % (
                       world, 'flat-bucket all-reduce in %d chunks started inside backward' % args.overlap_allreduce
                       if overlapped else 'one flat-bucket all-reduce'),
                   'gemm_pass_budget': budget or 'all 3xTF32',
                   'l2': 'working set >> L2: ~%d MB of [M,1140] fp32 activations per tensor, > 10 GB per step'
                         % (b * nmax * 1140 * 4 // 2 ** 20)},
        'e2e': {'value': total_graphs / (ms_e2e * 1e-3), 'unit': 'graphs/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': total_h2d // world, 'd2h_bytes_per_step': 4},
        'gpu_launches': launches,
        'clocks': clocks,
        'grad_allreduce_bytes': bucket.nbytes() if world > 1 else 0,
    }
    if prep_ms is not None:
        line['config']['prep_ms'] = prep_ms
        line['config']['prep'] = 'graph.sample_nodes (fuse / random) + graph.radius_graph on the GPU, once, untimed'
    if rank == 0 and not args.no_roofline:
        line.update(roofline_lines(args, device))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_graphs = cpu_sample_size(args, 4, 1, 25.0)
        rate, dt, cores, kind = cpu_reference_rate(args, 4, 1, n_graphs)
        what = ('verbatim reference model/network.py behind the PyG shim' if kind == 'reference'
                else 'dense CPU oracle port of the reference path')
        line['cpu_baseline'] = {'value': rate, 'unit': 'graphs/s', 'cores': cores, 'kind': kind,
                                'sample': '%d of the %d graphs per step (%s, fwd+bwd+Adam), 4 timed steps after 1 '
                                          'warm-up' % (n_graphs, b, what)}
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
