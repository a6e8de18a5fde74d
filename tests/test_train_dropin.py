"""The reference's own ``train.py`` running UNCHANGED on top of the B200 module.

CPU-only and only where ``/root/reference`` is mounted (it never travels to the GPU box): ``cgc_net_b200.compat``
stands in for torch_geometric / tensorboardX and maps ``from model import network`` to ``cgc_net_b200.network``;
the per-operator CPU formulas stand in for the CUDA library (fixture ``torch_backend``), so what is exercised is
the host-side contract: constructor call of train.py:254-261, ``DataParallel(model)(list_of_Data)``,
``(logits, loss)`` in training mode, optimiser / scheduler wiring, checkpoint format of common/utils.py:82-94."""
import os
import sys
import types

import pytest
import torch

from cgc_net_b200 import compat, graph

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'train.py')),
                                reason='reference checkout not mounted')


@pytest.fixture
def ref_train(torch_backend):
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.split('.')[0] in ('torch_geometric', 'tensorboardX', 'model', 'train', 'common', 'dataflow', 'setting')}
    for k in saved:
        sys.modules.pop(k, None)
    compat.install(replace_model=True, force=True)
    sys.path.insert(0, REF)
    argv = sys.argv
    sys.argv = ['train.py', '--load_data_list', '--input_feature_dim', '16', '--assign-ratio', '0.1', '--step_size', '1']
    try:
        import train as ref_train_module                     # the reference file, unmodified
        yield ref_train_module
    finally:
        sys.argv = argv
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k.split('.')[0] in ('torch_geometric', 'tensorboardX', 'model', 'train', 'common', 'dataflow', 'setting'):
                sys.modules.pop(k, None)
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


class _FakeLoader(object):
    """What prepare_train_val_loader returns, reduced to what train() touches: iterable of lists of Data with
    ``.dataset.set_epoch``."""

    def __init__(self, batches):
        self.batches = batches
        self.dataset = types.SimpleNamespace(set_epoch=lambda e: None, set_val_epoch=lambda e: None)

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def _data_list(n_graphs, seed):
    out = []
    for i in range(n_graphs):
        d = graph.synthetic_cell_graph(30 + 7 * i, 5, 16, seed=seed * 10 + i, row_is_centre=False)
        out.append(compat.Data(x=d.x, edge_index=d.edge_index, y=d.y, pos=d.pos, patch_idx=torch.tensor([i])))
    return out


def test_reference_train_loop_runs_unchanged(ref_train, tmp_path):
    from model import network                                  # train.py:16 resolves to the B200 module
    import cgc_net_b200.network as ours
    assert network is ours
    args = ref_train.arg_parse()
    args.num_epochs, args.resultdir = 2, str(tmp_path)
    args.num_classes = 3
    torch.manual_seed(0)
    # constructor call copied from train.py:254-261 (cell_graph); max_num_nodes shrunk so the CPU formulas are quick
    model = network.SoftPoolingGcnEncoder(300, args.input_feature_dim, args.hidden_dim, args.output_dim, True, True,
                                          args.hidden_dim, args.num_classes, args.assign_ratio, [50], concat=True,
                                          gcn_name=args.gcn_name, collect_assign=args.visualization,
                                          load_data_sparse=(args.load_data_list and not args.visualization),
                                          norm_adj=args.norm_adj, activation=args.activation, drop_out=args.drop_out,
                                          jk=args.jump_knowledge)
    from torch_geometric.nn import DataParallel              # train.py:17
    model = DataParallel(model)                                # train.py:285 (minus .cuda(): no GPU here)
    before = {k: v.clone() for k, v in model.module.state_dict().items()}
    loader = _FakeLoader([_data_list(3, 1), _data_list(2, 2)])
    trained, val_accs = ref_train.train(loader, model, args, val_dataset=None, test_dataset=None, writer=None)
    after = trained.module.state_dict()
    assert val_accs == []
    assert any(not torch.equal(before[k], after[k]) for k in before if 'weight' in k)       # Adam stepped
    assert int(after['GCN_embed_1.bn1.num_batches_tracked']) == 4                            # 2 epochs x 2 batches
    assert all(torch.isfinite(v).all() for v in after.values() if v.is_floating_point())


def test_checkpoint_round_trip_with_reference_format(ref_train, tmp_path):
    """save_checkpoint / load_checkpoint of common/utils.py:82-94 with the B200 module's state_dict; the file
    loads into the VERBATIM reference network (parameter names and shapes are the reference's own)."""
    from common.utils import load_checkpoint, save_checkpoint
    import cgc_net_b200.network as ours
    torch.manual_seed(3)
    args = (300, 16, 20, 20, True, True, 20, 3, 0.1, [50])
    m = ours.SoftPoolingGcnEncoder(*args, concat=True, load_data_sparse=True)
    path = os.path.join(str(tmp_path), 'run', 'weight.pth.tar')
    save_checkpoint({'epoch': 1, 'loss': 0.0, 'state_dict': m.state_dict(), 'optimizer': {}, 'val_acc': 0.0},
                    True, path)
    assert os.path.isfile(os.path.join(str(tmp_path), 'run', 'model_best.pth.tar'))
    ckpt = load_checkpoint(path)
    m2 = ours.SoftPoolingGcnEncoder(*args, concat=True, load_data_sparse=True)
    m2.load_state_dict(ckpt['state_dict'])                     # train.py:274
    # the verbatim reference model (oracle loader: its own torch_geometric shim) accepts the same file
    for k in list(sys.modules):
        if k.split('.')[0] in ('torch_geometric', 'model'):
            sys.modules.pop(k, None)
    from oracle import reference_loader
    ref_model = reference_loader.build_reference_encoder(*args, concat=True, load_data_sparse=True)
    ref_model.load_state_dict(ckpt['state_dict'])
    batch = graph.Batch.from_data_list([graph.synthetic_cell_graph(40, 5, 16, seed=9),
                                        graph.synthetic_cell_graph(25, 4, 16, seed=10)])
    m2.eval(), ref_model.eval()
    with torch.no_grad():
        a, b = m2(batch), ref_model(batch)
    assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max())
