"""Host-side logic that needs no GPU: state_dict compatibility with the reference, weight repacking,
the ConvLSTM gate-row interleave, batch sharding, and the world_size-2 (gloo) replica sync."""
import os
import types

import pytest
import torch
import torch.nn.functional as F

import recurrent_flows_msc_b200 as rf
from conftest import load_golden


def test_state_dict_keys_match_reference_fixtures():
    g = load_golden("listglow_cond")
    m = rf.ListGlow(g["x_size"], g["cond_sizes"], g["base_size"], types.SimpleNamespace(**g["args"]))
    assert set(m.state_dict().keys()) == set(g["sd"].keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(g["sd"][k].shape), k
    m.load_state_dict(g["sd"])
    g2 = load_golden("glowstep")
    a = types.SimpleNamespace(LU_decomposed=True, n_units_affine=16, non_lin_glow="relu", clamp_type="realnvp",
                              flow_norm="actnorm", flow_batchnorm_momentum=0.0)
    assert set(rf.GlowStep([2, 8, 4, 4], [2, 3, 4, 4], a).state_dict().keys()) == set(g2["sd"].keys())
    g3 = load_golden("invconv_plain")
    assert set(rf.InvConv(6, False).state_dict().keys()) == set(g3["sd"].keys())
    g4 = load_golden("convlstm")
    assert set(rf.ConvLSTM(3, 4, [3, 3]).state_dict().keys()) == set(g4["sd"].keys())
    g5 = load_golden("split2d_cond_softplus")
    assert set(rf.Split2d([2, 8, 4, 4], [2, 6, 4, 4]).state_dict().keys()) == set(g5["sd"].keys())


def test_invconv_init_is_orthogonal_lu():
    torch.manual_seed(0)
    m = rf.InvConv(12, True)
    W, W_inv, per_pixel = m.matrices()
    assert (W.t() @ W - torch.eye(12)).abs().max() < 1e-5        # SURVEY section 4 probe
    assert (W @ W_inv - torch.eye(12)).abs().max() < 1e-5
    assert abs(float(per_pixel)) < 1e-4
    w2, dl = m.get_weight(torch.zeros(2, 12, 3, 5), reverse=True)
    assert w2.shape == (12, 12, 1, 1) and abs(float(dl)) < 1e-3


def test_bad_enums_assert_like_reference():
    with pytest.raises(AssertionError):
        rf.Split2d([2, 8, 4, 4], [2, 6, 4, 4], clamp_function="tanh")
    with pytest.raises(AssertionError):
        rf.ActFun("gelu")


def test_batchnorm_options_state_dict():
    g = load_golden("listglow_batchnorm")
    m = rf.ListGlow(g["x_size"], g["cond_sizes"], g["base_size"], types.SimpleNamespace(**g["args"]))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in g["sd"].items()}
    m.load_state_dict(g["sd"])
    assert isinstance(m.glow_frame[1].norm, rf.Flow.BatchNormFlow)


def im2col_gemm(x, wp, cin_pad, N, k):
    """Reference of the kernel's GEMM view: A[pixel, tap*cin_pad + c] * Wp^T with zero 'same' padding."""
    B, C, H, W = x.shape
    p = (k - 1) // 2
    xp = F.pad(x, (p, p, p, p))
    cols = []
    for ky in range(k):
        for kx in range(k):
            t = xp[:, :, ky:ky + H, kx:kx + W]
            cols.append(F.pad(t, (0, 0, 0, 0, 0, cin_pad - C)))
    A = torch.cat(cols, 1).permute(0, 2, 3, 1).reshape(B * H * W, k * k * cin_pad)
    out = A @ wp.float().t()
    return out[:, :N].reshape(B, H, W, N).permute(0, 3, 1, 2)


@pytest.mark.parametrize("k", [1, 3])
def test_pack_conv_weight_layout(k):
    torch.manual_seed(1)
    x = torch.randn(2, 5, 4, 6).to(torch.bfloat16).float()
    w = torch.randn(7, 5, k, k).to(torch.bfloat16).float()
    wp, cin_pad = rf.ops.pack_conv_weight(w)
    assert wp.shape == (16, k * k * 32) and cin_pad == 32 and wp.dtype == torch.bfloat16
    torch.testing.assert_close(im2col_gemm(x, wp, cin_pad, 7, k), F.conv2d(x, w, None, 1, (k - 1) // 2), rtol=1e-4, atol=1e-4)
    perm = torch.tensor([3, 4, 0, 1, 2])   # staging order [cond | z1] vs the reference's cat[z1, cond]
    wp2, _ = rf.ops.pack_conv_weight(w, perm)
    torch.testing.assert_close(im2col_gemm(x[:, perm], wp2, cin_pad, 7, k), F.conv2d(x, w, None, 1, (k - 1) // 2), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("m_tiles", [1, 1 << 30])
@pytest.mark.parametrize("hidden", [4, 60, 64, 200, 256])
def test_lstm_gate_row_interleave(hidden, m_tiles):
    torch.manual_seed(2)
    cell = rf.ConvLSTMLayer(3, hidden, [3, 3], True)
    wgt, cin_pad, b, ht, ht_pad = cell._weights(m_tiles)
    if m_tiles == 1 and hidden == 200:
        assert ht == 8          # RFN: 25 narrow tiles instead of 5 wide ones
    assert hidden % ht == 0 and 4 * ht_pad <= 256 and ht_pad % 8 == 0
    n_tiles = hidden // ht
    assert wgt.shape[0] == n_tiles * 4 * ht_pad
    w = cell.conv[0].weight.detach()
    ref, _ = rf.ops.pack_conv_weight(w)
    for t in range(n_tiles):
        for gate in range(4):
            rows = slice((t * 4 + gate) * ht_pad, (t * 4 + gate) * ht_pad + ht)
            src = slice(gate * hidden + t * ht, gate * hidden + t * ht + ht)
            assert torch.equal(wgt[rows], ref[src])
            assert torch.equal(b[rows], cell.conv[0].bias.detach()[src])
            assert float(wgt[(t * 4 + gate) * ht_pad + ht:(t * 4 + gate + 1) * ht_pad].float().abs().sum()) == 0


def test_versioned_cache_tracks_inplace_updates():
    m = rf.Conv2dZeros(4, 6)
    s0, _ = m.affine()
    with torch.no_grad():
        m.logs.add_(1.0)
    s1, _ = m.affine()
    assert not torch.equal(s0, s1)
    assert m.affine()[0] is s1


def test_shard_range_covers_batch():
    for n in (0, 1, 7, 30, 570):
        for ws in (1, 2, 3, 8):
            spans = [rf.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    x = torch.arange(10)
    assert torch.equal(torch.cat([rf.shard_batch(x, r, 3) for r in range(3)]), x)
    a, b = rf.shard_batch([x, [x, None]], 1, 2)
    assert torch.equal(a, x[5:]) and b[1] is None


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)      # replicas start different (as after a per-shard ActNorm init)
        a = types.SimpleNamespace(LU_decomposed=True, n_units_affine=8, non_lin_glow="relu", clamp_type="realnvp",
                                  flow_norm="actnorm", flow_batchnorm_momentum=0.0)
        m = rf.GlowStep([2, 4, 4, 4], [2, 2, 4, 4], a)
        with torch.no_grad():
            m.norm.logs.normal_()
        n = rf.sync_module_state(m, src=0)
        flat = torch.cat([t.detach().float().reshape(-1) for t in list(m.parameters()) + list(m.buffers())])
        got = rf.parallel.gather_batch(flat[None])
        lo, hi = rf.shard_range(7, rank, world)
        q.put((rank, n, bool(torch.equal(got[0], got[1])), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sync_and_shards():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(same for _, _, same, _ in res)
    assert res[0][1] > 0 and res[0][3] == (0, 4) and res[1][3] == (4, 7)


def test_invalidate_caches_epoch():
    """Raw-pointer parameter updates (fused Adam, CUDA-graph replay) do not bump torch's version counters; the caches
    must then be dropped through invalidate_caches()."""
    from recurrent_flows_msc_b200.Flow.glow_modules import _Versioned, invalidate_caches
    p = torch.nn.Parameter(torch.ones(3))
    c = _Versioned()
    calls = []

    def build():
        calls.append(1)
        return p.detach().clone()
    a = c.get("k", (p,), build)
    assert c.get("k", (p,), build) is a and len(calls) == 1
    p.data.view(-1)[0] = 5.0            # a write torch does not see (what a kernel writing through data_ptr() does)
    assert c.get("k", (p,), build) is a
    invalidate_caches()
    b = c.get("k", (p,), build)
    assert len(calls) == 2 and float(b[0]) == 5.0


def test_zero_arena_hands_out_disjoint_zeroed_views():
    from recurrent_flows_msc_b200 import ops
    dev = torch.device("cpu")
    with ops.zero_arena(dev):
        a = ops._zeros(10, dev)
        b = ops._zeros(7, dev)
        a.add_(1.0)
        assert float(b.abs().sum()) == 0.0 and a.data_ptr() != b.data_ptr()
        assert (b.data_ptr() - a.data_ptr()) % 16 == 0          # 4-float granularity keeps 16-byte alignment
        big = ops._zeros(ops._ZeroArena.CHUNK + 5, dev)           # larger than a chunk: gets its own allocation
        assert big.numel() == ops._ZeroArena.CHUNK + 5 and float(big.abs().sum()) == 0.0
    c = ops._zeros(3, dev)                                       # outside the context: a plain allocation
    assert c.shape == (3,) and float(c.abs().sum()) == 0.0


def test_training_entry_points_need_cuda():
    """The training path has no CPU fallback either: log_prob / ConvLSTM under autograd on CPU tensors raise."""
    import types
    import pytest
    import recurrent_flows_msc_b200 as rf
    a = types.SimpleNamespace(LU_decomposed=True, n_units_affine=16, non_lin_glow="relu", clamp_type="realnvp",
                              flow_norm="actnorm", flow_batchnorm_momentum=0.0, learn_prior=False, n_units_prior=16,
                              make_conditional=False, base_norm="actnorm", split2d_act="softplus", L=1, K=1, n_bits=8)
    m = rf.ListGlow([2, 1, 8, 8], [[2, 0, 4, 4]], [2, 0, 4, 4], a)
    with pytest.raises(Exception):
        m.log_prob(torch.zeros(2, 1, 8, 8), [torch.zeros(2, 0, 4, 4)], None)
    lstm = rf.ConvLSTM(4, 4, [3, 3])
    with pytest.raises(Exception):
        lstm(torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(RuntimeError):
        rf.FlatAdam(lstm.parameters())


def test_clear_workspaces_refuses_while_a_graph_is_alive():
    """a captured graph holds raw pointers into the workspace pool."""
    import recurrent_flows_msc_b200 as rf

    class FakeGraph:
        pass
    g = FakeGraph()
    rf.ops.register_graph(g)
    with pytest.raises(rf._lib.RfkError):
        rf.ops.clear_workspaces()
    del g
    import gc
    gc.collect()
    rf.ops.clear_workspaces()


def test_split_k_heuristic():
    """ops.choose_k_split: nine slices for a single pixel tile with a long K (sampling, deepest level), three filter rows for
    16..24 pixel tiles (deepest level of the 570-frame workload), none for short K, 1x1 convs, narrow inputs or many tiles."""
    from recurrent_flows_msc_b200 import ops
    if ops.SPLIT:
        pytest.skip("split-precision mode never splits K")
    assert ops.choose_k_split(120, 9, 320) == 9            # 30 sequences at 2x2: one tile, K = 2880
    assert ops.choose_k_split(2280, 9, 320) == 3           # 570 frames at 2x2: 18 tiles
    assert ops.choose_k_split(2280, 9, 128) == 1           # K = 1152: too short to pay for the reduction
    assert ops.choose_k_split(9120, 9, 192) == 1           # 72 tiles already cover the GPU
    assert ops.choose_k_split(1920, 9, 128) == 1           # 15 tiles (sampling, 8x8 maps): measured slower when split
    assert ops.choose_k_split(2280, 1, 320) == 1 and ops.choose_k_split(2280, 9, 32) == 1


def test_gemm_m_tiles_matches_the_kernel_tiling():
    """ops.gemm_m_tiles mirrors the conv kernels' tile = NIMG x TH x TW pixels with power-of-two TW, TH (csrc/conv_gemm.cu
    m_tiles_of): the thresholds of the one-kernel coupling network are expressed in these tiles."""
    from recurrent_flows_msc_b200 import ops
    assert ops.gemm_m_tiles(570, 32, 32) == 4560 and ops.gemm_m_tiles(570, 16, 16) == 1140
    assert ops.gemm_m_tiles(30, 32, 32) == 240 and ops.gemm_m_tiles(30, 16, 16) == 60
    assert ops.gemm_m_tiles(30, 2, 2) == 1 and ops.gemm_m_tiles(570, 2, 2) == 18
    assert ops.gemm_m_tiles(7, 12, 20) == 21               # ragged: tiles of 32 x 4 pixels, three per image
    assert ops.gemm_m_tiles(1, 200, 200) == 2 * 200        # tiles of 128 x 1 pixels


def test_one_kernel_coupling_network_is_gated():
    """The module-level switch to rfk_coupling_nn_fused needs settled ActNorms (foldable), at least FUSE_NN_MIN_TILES pixel
    tiles and at most FUSE_NN_MAX_PLANES tap planes; recompute mode is an attribute / environment default."""
    import recurrent_flows_msc_b200 as rf
    from recurrent_flows_msc_b200.Flow import glow_modules as gm, training
    assert gm.FUSE_NN_MIN_TILES == int(os.environ.get("RFK_FUSE_NN_MIN_TILES", "48"))
    assert gm.FUSE_NN_MAX_PLANES == int(os.environ.get("RFK_FUSE_NN_MAX_PLANES", "128"))
    layer = rf.Flow.Conv2dNorm(18, 64)
    assert not layer.foldable()                            # ActNorm not initialised (and not on a GPU)
    layer.norm_type.mark_initialized()
    assert layer.norm_type.is_initialized() and not layer.foldable()   # CPU parameters: the packing kernel needs CUDA
    assert training.RECOMPUTE == (os.environ.get("RFK_RECOMPUTE", "0") == "1")
