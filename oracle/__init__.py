"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds the CPU checker for the stochastic-aggregation hot path of
yuanqing-wang/stag: a torch/numpy restatement of the reference algorithm
(ref_index.py, ref_spmm.py, ref_layers.py), a numpy restatement of the library's own
counter-based RNG (ref_philox.py), a plain-C restatement used as the CPU baseline
(csrc/), and a pure-PyTorch DGL shim (dgl_shim/) that lets the reference's own
Python code run unmodified in the build container.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import or execute anything here.  The product
(``stag_b200/``) never does, and fails loudly when its CUDA library is missing.
"""
