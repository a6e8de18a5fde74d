"""Import ``/root/reference/model/network.py`` VERBATIM behind the PyG shim.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Only usable where the read-only
reference checkout exists (the build container); the GPU box has no ``/root/reference``,
unless a copy has been staged under ``baseline/_ref`` (``oracle/stage_reference.py``) -- ``available()`` lets callers
skip or fall back to the restatement.

The only patch applied to the imported module is ``construct_mask``: the reference
hard-codes ``.cuda()`` at ``model/network.py:180``; on a CPU-only host the tensor stays
where it is.  Nothing else is touched.
"""
import importlib
import os
import sys
import warnings

import torch

from . import pyg_shim

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """CGC_REFERENCE_ROOT, the read-only checkout of the build container, or the staged copy that travels to the
    GPU box (oracle/stage_reference.py -> baseline/_ref, git-ignored)."""
    for cand in (os.environ.get('CGC_REFERENCE_ROOT'), '/root/reference', os.path.join(_REPO, 'baseline', '_ref')):
        if cand and os.path.isfile(os.path.join(cand, 'model', 'network.py')):
            return cand
    return '/root/reference'


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'model', 'network.py'))


_network = None


def load_network_module(cpu=False):
    """Return the verbatim ``model.network`` module of the reference.  ``cpu=True``: the model will run on host
    tensors although a GPU is present (the CPU arm of bench.py on the GPU box), so ``construct_mask`` must not
    move its result to the device either."""
    global _network
    if _network is not None:
        return _network
    if not available():
        raise FileNotFoundError('reference checkout not found at %s' % REFERENCE_ROOT)
    pyg_shim.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module('model.network')

    if cpu or not torch.cuda.is_available():
        orig = mod.SoftPoolingGcnEncoder.construct_mask

        def construct_mask_cpu(max_nodes, batch_num_nodes):
            # same arithmetic as model/network.py:172-180 without the trailing .cuda()
            out = torch.zeros(len(batch_num_nodes), max_nodes)
            for i, n in enumerate(batch_num_nodes):
                out[i, :int(n)] = 1
            return out.unsqueeze(2)

        construct_mask_cpu.__wrapped__ = orig
        mod.SoftPoolingGcnEncoder.construct_mask = staticmethod(construct_mask_cpu)
    _network = mod
    return mod


def build_reference_encoder(*args, cpu=False, **kwargs):
    """``model.network.SoftPoolingGcnEncoder(*args, **kwargs)`` of the verbatim reference."""
    mod = load_network_module(cpu=cpu)
    return mod.SoftPoolingGcnEncoder(*args, **kwargs)


def run_reference(model, data):
    """Forward with deprecation noise (``size_average``, network.py:289) silenced."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return model(data)
