"""CPU oracle for the CGC-Net graph-convolution hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or as the thing timed on the host
cores as the reference arm) -- never as a fallback for the CUDA path.

Contents
--------
``pyg_shim``       restatement of the four ``torch_geometric==1.2.1`` symbols the
                   reference imports (the package is not vendored in /root/reference and
                   not installable offline); lets ``/root/reference/model/network.py`` be
                   imported VERBATIM in the build container.
``reference_loader`` imports the verbatim reference behind the shim (only where
                   ``/root/reference`` exists, i.e. the build container).
``restatement``    self-contained dense pure-torch restatement of
                   ``model/network.py`` + ``model/utils.py`` (this is what travels to
                   the GPU box); checked against the verbatim import and against the
                   committed golden vectors in ``tests/golden``.
``op_reference``   per-operator CPU formulas (forward and hand-derived backward) of every
                   C-ABI entry point, used by the tests as the op-level checker.
``make_golden``    script that minted ``tests/golden/*.npz`` from the verbatim reference.

Parity pin: the reference ships no golden vectors or tests, so the
oracle is pinned against outputs of the reference itself run in the build container
(``make_golden.py``), with the PyG layer restated from the pinned upstream versions
(``requirements.txt:38-42``).  At the PyG boundary parity is therefore pinned to the
restated semantics, not to upstream binaries.
"""
