"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the Polyphemus message-passing hot path used as the parity checker:

* ``graph_oracle``  — numpy restatement of ``data.py:14-204`` (per-bar typed edges, collation).
* ``model_oracle``  — functional fp32/fp64 PyTorch-CPU restatement of ``model.py:30-135,167-208``
                      (GCL/GCN) and of the rest of the VAE + ``training.py:298-347`` losses.
* ``pyg_shim`` / ``ref_loader`` — load the reference's *own unmodified files* from
  ``/root/reference`` on top of a stand-in for the un-vendored ``torch-geometric==2.0.2`` symbols.
  Only usable where ``/root/reference`` exists (the build container); used to pin the restatement
  and to generate ``tests/golden/*`` (see ``tests/golden/make_golden.py``).

Parity status: the reference ships NO tests or golden vectors for this path.
The restatement is pinned against outputs of the reference's own code executed in the build
container (committed under ``tests/golden/``); the PyG semantics themselves are recalled, not
verified against a real PyG install ("parity pinned to reference-on-shim").

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package. The product (``polyphemus_b200``) never does.
"""
