"""CPU: the C-ABI library loads and exports every symbol include/stag_b200.h declares; the
host-side mirror of the reference interface (constructors, state_dict keys, error behaviour);
no compute without a GPU -- and no CPU fallback."""
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "stag_b200.h")).read()
    return sorted(set(re.findall(r"STAG_API\s+[\w\s\*]+?\b(stag_\w+)\s*\(", text)))


def test_library_exports_every_header_symbol(lib):
    from stag_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 12, syms
    for s in syms:
        assert hasattr(lib, s), "header declares %s but the library does not export it" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding and header disagree"
    assert lib.stag_abi_version() == 2
    assert lib.stag_hub_threshold() > 0 and lib.stag_hub_segment() > 0


def test_struct_layout_matches_header():
    import ctypes
    from stag_b200 import _lib
    assert ctypes.sizeof(_lib.StagGraph) == 112
    assert ctypes.sizeof(_lib.StagNoise) == 72


def test_argument_errors_are_reported_not_thrown(lib):
    from stag_b200 import _lib
    import ctypes
    g = _lib.StagGraph()
    n = _lib.StagNoise()
    rc = lib.stag_spmm_fwd(ctypes.byref(g), 0, 4, 0, 4, 1, ctypes.byref(n), 0, 0, 0, 4, 0, 0, 0, 0, 0)
    assert rc == _lib.STAG_EINVAL
    assert lib.stag_last_error()
    assert lib.stag_noise_emit(ctypes.byref(n), 10, 1, 0, 0, 0) == _lib.STAG_EINVAL


def test_no_cpu_fallback():
    import stag_b200 as sb
    g = sb.rand_graph(5, 20)
    with pytest.raises(sb.StagLibraryError):
        sb.ops.stochastic_aggregate(g, torch.randn(5, 8), None)
    layer = sb.layers.StagLayer(sb.zoo.GCN(8, 4))
    with pytest.raises(sb.StagLibraryError):
        layer(g, torch.randn(5, 8))


def test_missing_library_fails_loudly(tmp_path):
    from stag_b200 import _lib
    with pytest.raises(_lib.StagLibraryError):
        _lib.load(str(tmp_path / "nope.so"))


def test_state_dict_keys_match_reference():
    import stag_b200 as stag
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32))
    assert list(layer.state_dict()) == ["base_layer.weight", "base_layer.bias", "q_a.loc", "q_a.scale",
                                        "p_a.loc", "p_a.scale"]
    q = torch.distributions.Normal(torch.ones(16), torch.ones(16))
    layer = stag.layers.StagLayer(stag.zoo.GCN(16, 32), q_a=q, vi=True)
    assert list(layer.state_dict()) == ["base_layer.weight", "base_layer.bias", "q_a.loc", "q_a.log_scale",
                                        "p_a.loc", "p_a.log_scale"]
    assert isinstance(layer.q_a.log_scale, torch.nn.Parameter) and layer.q_a.loc.shape == (16,)
    assert layer.kl_divergence().item() == pytest.approx(0.0, abs=1e-7)
    assert stag.layers.StagLayer(stag.zoo.GCN(4, 4)).kl_divergence() == 0.0
    sage = stag.zoo.GraphSAGE(8, 4)
    assert sorted(sage.state_dict()) == ["bias", "fc_neigh.weight", "fc_self.weight"]
    assert sorted(stag.zoo.GraphSAGE(8, 4, aggregator_type="gcn").state_dict()) == ["bias", "fc_neigh.weight"]
    assert stag.zoo.GAT(8, 4).sample_dimension == 4
    assert sorted(stag.zoo.GIN(8, 4).state_dict()) == ["apply_func.bias", "apply_func.weight", "eps"]


def test_distributions_interface():
    """stag/tests/test_distributions.py:8-46."""
    import stag_b200 as stag
    D = stag.distributions
    d = D.ParametrizedDistribution(torch.distributions.Normal(0.0, 1.0))
    assert d.expand([10, 8]).rsample().shape == (10, 8)
    d = D.ParametrizedDistribution(torch.distributions.Normal(0.0, 1.0), vi=True)
    assert sorted(n for n, _ in d.named_parameters()) == ["loc", "log_scale"]
    d = D.ParametrizedDistribution(torch.distributions.Normal(torch.zeros(10, 8), torch.ones(10, 8)))
    assert d.expand([12, 11, 10, 8]).rsample().shape == (12, 11, 10, 8)
    assert D.DeltaDistribution(0.0).sample() == 0.0
    a = D.AmortizedDistribution(16, 1)
    assert a.new_parameter_names == ["loc", "log_scale"]
    assert d.fused_parameters()[0] == "normal"
    assert D.ParametrizedDistribution(torch.distributions.Bernoulli(probs=0.3)).fused_parameters()[0] == "bernoulli"
    assert D.ParametrizedDistribution(torch.distributions.Uniform(0.0, 2.0)).fused_parameters()[0] == "uniform"
    assert D.ParametrizedDistribution(torch.distributions.Laplace(0.0, 2.0)).fused_parameters() is None


def test_noise_spec_parameter_shape_classes():
    from stag_b200 import _lib
    from stag_b200.ops import NoiseSpec
    E, K = 30, 8
    t = torch.ones
    assert NoiseSpec("normal", t(()), t(()), K, E, seed=0, offset=0).param_shape == _lib.PARAM_SCALAR
    assert NoiseSpec("normal", t(K), t(K), K, E, seed=0, offset=0).param_shape == _lib.PARAM_CHANNEL
    assert NoiseSpec("normal", t(E, 1), t(E, 1), K, E, seed=0, offset=0).param_shape == _lib.PARAM_EDGE
    assert NoiseSpec("normal", t(E, K), t(E, K), K, E, seed=0, offset=0).param_shape == _lib.PARAM_EDGE_CHANNEL
    s = NoiseSpec("normal", t(()), t(K), K, E, seed=0, offset=0)   # mixed -> widened
    assert s.param_shape == _lib.PARAM_CHANNEL and s.p0.shape == (K,)
    with pytest.raises(ValueError):
        NoiseSpec("normal", t(5), t(5), K, E, seed=0, offset=0)
    with pytest.raises(ValueError):
        NoiseSpec("laplace", t(()), t(()), K, E, seed=0, offset=0)


def test_philox_call_counter():
    import stag_b200 as sb
    sb.manual_seed(5)
    a = sb.random.next_offset()
    b = sb.random.next_offset()
    assert a == (5, 0) and b == (5, 1)
    sb.manual_seed(5)
    assert sb.random.next_offset() == (5, 0)


def test_graph_helpers_match_dgl_semantics():
    import stag_b200 as sb
    g = sb.Graph(torch.tensor([0, 1, 1, 2]), torch.tensor([1, 1, 2, 0]), 4)
    assert g.number_of_nodes() == 4 and g.number_of_edges() == 4
    assert g.in_degrees().tolist() == [1, 2, 1, 0] and g.out_degrees().tolist() == [1, 2, 1, 0]
    s, d = sb.add_self_loop(g).edges()
    assert s.tolist() == [0, 1, 1, 2, 0, 1, 2, 3] and d.tolist() == [1, 1, 2, 0, 0, 1, 2, 3]
    s, d = sb.remove_self_loop(g).edges()
    assert s.tolist() == [0, 1, 2] and d.tolist() == [1, 2, 0]
    s, d = sb.add_reverse_edges(g).edges()
    assert s.tolist() == [0, 1, 1, 2, 1, 1, 2, 0]
    b = sb.batch([g, g])
    assert b.batch_num_nodes().tolist() == [4, 4] and b.edges()[0].tolist() == [0, 1, 1, 2, 4, 5, 5, 6]
    lv = g.local_var()
    lv.ndata["h"] = torch.zeros(4)
    assert "h" not in g.ndata
    with g.local_scope():
        g.ndata["x"] = torch.zeros(4)
    assert "x" not in g.ndata


def test_early_stopping_contract():
    import stag_b200 as sb
    es = sb.utils.EarlyStopping(patience=2)
    m = torch.nn.Linear(2, 2)
    assert es([1.0, 1.0], m) is False
    assert es([0.5, 0.5], m) is False and es.best_state is not None
    assert es([0.6, 0.6], m) is False
    assert es([0.7, 0.7], m) is True


@pytest.mark.parametrize("out_features,hidden", [(1, None), (16, None), (16, 8)])
def test_amortized_condition_equals_the_concatenated_form(out_features, hidden):
    """AmortizedDistribution.condition applies the first Linear to the node rows (split into its source / destination
    halves) instead of the [E, 2 D] concat of the reference (stag/distributions.py:221-233): same parameters."""
    import torch
    import stag_b200 as stag
    torch.manual_seed(0)
    N, E, D = 30, 200, 16
    g = stag.Graph(torch.randint(0, N, (E,)), torch.randint(0, N, (E,)), N)
    q = stag.distributions.AmortizedDistribution(D, out_features, hidden_features=hidden).double()
    for feat in (torch.randn(N, D, dtype=torch.float64), torch.randn(3, N, D, dtype=torch.float64)):
        q.condition(g, feat)
        src, dst = g.edges()
        h = q.embedding_mlp(torch.cat([feat.index_select(-2, src), feat.index_select(-2, dst)], dim=-1))
        for key in q.new_parameter_names:
            ref = q.parameters_mlp[key](h)
            assert q.new_parameters[key].shape == ref.shape
            assert torch.allclose(q.new_parameters[key], ref, rtol=1e-12, atol=1e-12)
        dist = q.base_distribution
        assert dist.loc.shape[-2:] == (E, out_features) and (dist.scale > 0).all()


def test_philox_stream_follows_torch_seed_and_is_process_global():
    """The default seed is torch's: torch.manual_seed restarts the stream, threads share one counter,
    fold_rank separates data-parallel ranks, get_state / set_state resume it."""
    import threading
    import stag_b200 as sb
    torch.manual_seed(1234)
    assert sb.random.next_offset() == (1234, 0) and sb.random.next_offset() == (1234, 1)
    torch.manual_seed(99)
    assert sb.random.next_offset() == (99, 0)
    got = []
    t = threading.Thread(target=lambda: got.append(sb.random.next_offset()))
    t.start(); t.join()
    assert got == [(99, 1)]                       # one process-global counter, not one per thread
    state = sb.random.get_state()
    assert state == (99, 2)
    sb.random.fold_rank(0)
    s0 = sb.random.get_state()[0]
    sb.random.set_state(*state)
    sb.random.fold_rank(1)
    s1 = sb.random.get_state()[0]
    assert s0 != s1 and s0 != 99 and sb.random.get_state()[1] == 0
    sb.random.set_state(*state)
    assert sb.random.next_offset() == (99, 2)
    sb.manual_seed(5)
    assert sb.random.next_offset() == (5, 0)


def test_model_api_carries_sample_base_and_layers_pickle():
    import copy
    import inspect
    import pickle
    import stag_b200 as sb
    for fn in (sb.models.StagModel.forward, sb.models.StagModel.loss, sb.models.StagModel.loss_terms):
        assert "sample_base" in inspect.signature(fn).parameters
    layer = sb.layers.StagLayer(sb.zoo.GCN(4, 3))
    g = sb.Graph(torch.tensor([0, 1]), torch.tensor([1, 0]), 2)
    layer._graph_ref = sb.layers._GraphRef(g)       # what a forward leaves behind
    layer._noise_tensor = torch.ones(2, 4)
    clone = copy.deepcopy(layer)
    assert clone._graph_ref is None and clone._noise_tensor is None
    assert pickle.loads(pickle.dumps(layer)).base_layer.weight.shape == (4, 3)
    assert layer._last_graph is not None
    del g
    assert layer._last_graph is None                 # weak: the layer does not keep graphs (and their CSC / CSR) alive
    # GatedGCN runs the samples sequentially (per-pass BatchNorm statistics), the other fused layers batch them
    assert sb.zoo.GCN.accepts_sample_batch and not sb.zoo.GatedGCN.accepts_sample_batch
    m = sb.models.StagModel(torch.nn.ModuleList([sb.layers.StagLayer(sb.zoo.GatedGCN(4, 4))]))
    assert m._can_batch() is False


def test_generator_selection_rules_of_a_noise_spec():
    """Which standard-normal generator a NoiseSpec resolves to (no device needed): the tensor-core one where the fused
    kernel takes it -- K a multiple of 128, or K >= 96 within a third of one on graphs of >= 2^18 edges (run
    zero-padded) -- with scalar / per-edge parameters without gradients and no relu / in-norm; Box-Muller otherwise;
    an explicit 'hadamard' outside that set is refused, never re-routed."""
    import pytest
    import torch
    from stag_b200 import _lib
    from stag_b200.ops import NoiseSpec
    one = torch.ones(())
    big = 1 << 18
    for K, E, width in [(128, 10, 128), (384, 10, 384), (100, big, 128), (100, big - 1, 0), (97, big, 128), (95, big, 0),
                        (130, big, 0), (200, big, 256), (1433, big, 1536), (64, big, 0), (1, big, 0)]:
        sp = NoiseSpec("normal", one, one, K, E)
        assert sp.hadamard_width == width, (K, E)
        assert sp.lib_kind == (_lib.NOISE_NORMAL_HADAMARD if width else _lib.NOISE_NORMAL)
        assert NoiseSpec("normal", one, one, K, E, generator="boxmuller").lib_kind == _lib.NOISE_NORMAL
        # every view of a spec agrees (the emitted tensor and the fused passes must draw the same stream)
        assert sp.with_samples(1).lib_kind == sp.with_samples(16, 3).lib_kind == sp.lib_kind
    for kw in (dict(relu=True), dict(in_norm=True)):
        assert NoiseSpec("normal", one, one, 128, 10, **kw).lib_kind == _lib.NOISE_NORMAL
        with pytest.raises(ValueError):
            NoiseSpec("normal", one, one, 128, 10, generator="hadamard", **kw).lib_kind
    learn = torch.ones(128, requires_grad=True)
    assert NoiseSpec("normal", learn, torch.ones(128), 128, 10).lib_kind == _lib.NOISE_NORMAL     # vi: two-sum kernel
    assert NoiseSpec("uniform", one, 2 * one, 128, 10).lib_kind == _lib.NOISE_UNIFORM
    with pytest.raises(ValueError):
        NoiseSpec("normal", one, one, 128, 10, generator="ziggurat")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU (it times the C/OpenMP restatement of the reference algorithm on the host
    cores): one JSON line with the arm's metric / unit / config and the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GEdge-samples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["gpu_launches"] == 0
