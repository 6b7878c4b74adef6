"""GPU tests of the tensor-core path (precision 'tf32': tcgen05.mma kind::tf32, TMA-fed, TMEM
accumulators).  Tolerance per north_star: activations / weights within 1e-3 of the fp32 reference
(TF32 keeps 10 mantissa bits of W and of real-valued activations; accumulation is fp32)."""
import ctypes as C

import pytest
import torch

from oracle import rbm_oracle as O
from oracle.philox import RandomField

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture()
def M():
    import multimodal_idbn_b200 as m
    m.load_library()
    m.set_precision("tf32")
    yield m
    m.set_precision("fp32")


def make(M, V, H, seed=0, scale=1.0, groups=None):
    st = O.new_state(V, H, seed=seed, groups=groups, lr=0.1, weight_decay=1e-4, momentum=0.5,
                     final_momentum=0.95, dynamic_lr=True)
    st.W *= scale
    g = torch.Generator().manual_seed(seed + 1)
    st.hb.copy_(torch.randn(H, generator=g) * 0.1); st.vb.copy_(torch.randn(V, generator=g) * 0.1)
    r = M.RBM(V, H, 0.1, 1e-4, 0.5, dynamic_lr=True, final_momentum=0.95,
              softmax_groups=list(groups or [])).to(DEV)
    with torch.no_grad():
        r.W.data.copy_(st.W); r.hid_bias.data.copy_(st.hb); r.vis_bias.data.copy_(st.vb)
    return st, r


SHAPES = [(10000, 1500), (1500, 500), (532, 256), (128, 128), (36, 20)]


@pytest.mark.parametrize("V,H", SHAPES)
@pytest.mark.parametrize("B", [1, 64, 130, 256, 700])
def test_up_down_tf32_vs_oracle(M, V, H, B):
    st, r = make(M, V, H, seed=V + H)
    gen = torch.Generator().manual_seed(B)
    v = (torch.rand(B, V, generator=gen) < 0.3).float()
    h = torch.rand(B, H, generator=gen)
    p = r.forward(v.to(DEV))
    torch.testing.assert_close(p.cpu(), O.hidden_probs(st, v), rtol=0, atol=1e-3)
    pv = r.visible_probs(h.to(DEV))
    torch.testing.assert_close(pv.cpu(), O.visible_probs(st, h), rtol=0, atol=1e-3)
    # real-valued (non-binary) inputs, as test_extraction.py feeds (randn)
    x = torch.randn(B, V, generator=gen)
    torch.testing.assert_close(r.forward(x.to(DEV)).cpu(), O.hidden_probs(st, x), rtol=0, atol=2e-3)


@pytest.mark.parametrize("V,H", SHAPES)
@pytest.mark.parametrize("B", [1, 64, 100])
def test_assoc_stats_tf32(M, V, H, B):
    from multimodal_idbn_b200 import _lib as L
    st, r = make(M, V, H, seed=3)
    gen = torch.Generator().manual_seed(7)
    vp = (torch.rand(B, V, generator=gen) < 0.3).float(); vn = (torch.rand(B, V, generator=gen) < 0.3).float()
    hp = torch.rand(B, H, generator=gen); hn = torch.rand(B, H, generator=gen)
    ref = vp.T @ hp - vn.T @ hn
    out = torch.empty(V, H, device=DEV)
    ctx, stream = L.context_for(out)
    rs = r._struct()
    d = [t.to(DEV).contiguous() for t in (vp, hp, vn, hn)]
    ctx.check(ctx.lib.imdbn_assoc_stats(ctx.handle, C.byref(rs), L.ptr(d[0]), L.ptr(d[1]), L.ptr(d[2]),
                                        L.ptr(d[3]), B, L.ptr(out), stream), "assoc")
    # kind::tf32 reads the top 19 bits of every fp32 operand (truncation): emulate that exactly
    def trunc(t):
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref_t = trunc(vp).double().T @ trunc(hp).double() - trunc(vn).double().T @ trunc(hn).double()
    if B <= 64:      # larger batches take the exact fp32 engine for this product
        torch.testing.assert_close(out.cpu().double(), ref_t, rtol=1e-5, atol=2e-5)
    # and against the un-truncated product the error is the TF32 input rounding, ~2^-11 per term
    assert float((out.cpu() - ref).abs().max()) < 2e-3 * max(4.0, 0.3 * B) ** 0.5 + 1e-3 * 0.3 * B


def test_cd1_update_tf32_c1_shape(M):
    V, H, B = 10000, 1500, 64
    st, r = make(M, V, H, seed=11)
    data = O.synthetic_images(B, V, seed=1234)
    W0 = st.W.clone()
    loss_ref, _ = O.cd_train(st, data, 0, 1, RandomField(5, 0))
    r.set_rng(5, 0)
    loss = r.train_epoch(data.to(DEV), 0, 1, CD=1)
    torch.testing.assert_close(loss.cpu(), loss_ref, rtol=1e-3, atol=1e-6)
    dW_ref = st.W - W0
    dW = r.W.detach().cpu() - W0
    # a few sampled units flip under TF32 (each flip moves one row / column of dS by lr/B); the bulk of
    # the update must agree to 1e-3 of its scale
    scale = float(dW_ref.abs().mean())
    assert float((dW - dW_ref).abs().mean()) < 0.05 * scale
    assert float(((dW - dW_ref).abs() > 1e-3 * 0.1).float().mean()) < 0.02
    torch.testing.assert_close(r.hid_bias.detach().cpu(), st.hb, rtol=0, atol=2e-3)


def test_tf32_and_fp32_modes_agree_on_a_training_run(M):
    """50 CD-1 updates of a 532->256 joint-shaped RBM in both modes, same random field: the
    reconstruction error curves must track each other (north_star: statistically matched)."""
    V, H, B = 532, 256, 64
    losses = {}
    for mode in ("fp32", "tf32"):
        M.set_precision(mode)
        _, r = make(M, V, H, seed=5, scale=1.0)
        r.set_rng(77, 0)
        data = O.synthetic_images(B * 4, V, p=0.2, seed=3).to(DEV)
        ls = [r.train_epoch(data[(i % 4) * B:(i % 4 + 1) * B], 0, 1, CD=1) for i in range(50)]
        losses[mode] = torch.stack(ls).cpu()
    M.set_precision("tf32")
    assert losses["tf32"][-1] < losses["tf32"][0]
    torch.testing.assert_close(losses["tf32"], losses["fp32"], rtol=2e-2, atol=1e-3)


def test_large_batch_chains_on_tensor_cores(M):
    """B >= 512 mean-field chains run step by step on the tcgen05 passes (chain_stepped.cuh)."""
    V, H, Dz, K, B = 532, 256, 500, 32, 600
    st, r = make(M, V, H, seed=21, scale=2.0, groups=[(Dz, V)])
    z = torch.rand(B, Dz, generator=torch.Generator().manual_seed(2))
    y = O.synthetic_labels(B, K, seed=3)
    mu = torch.rand(B, Dz, generator=torch.Generator().manual_seed(4))
    vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, Dz:] = y; km[:, Dz:] = 1
    ref = O.noisy_meanfield(st, vk, km, n_steps=20, mu_pull=(mu, 0.15), fld=RandomField(31, 0))
    r._mu_pull = {"mu_k": mu.to(DEV), "eta0": 0.15}
    r.set_rng(31, 0)
    out = r.noisy_meanfield_annealed(vk.to(DEV), km.to(DEV), n_steps=20)
    r._mu_pull = None
    torch.testing.assert_close(out.cpu(), ref, rtol=0, atol=3e-3)
    vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, :Dz] = z; km[:, :Dz] = 1
    ref = O.conditional_gibbs(st, vk, km, n_steps=20, fld=RandomField(31, 1))
    r.set_rng(31, 1)
    out = r.conditional_gibbs(vk.to(DEV), km.to(DEV), n_steps=20)
    torch.testing.assert_close(out.cpu(), ref, rtol=0, atol=3e-3)
    assert torch.allclose(out[:, Dz:].sum(1).cpu(), torch.ones(B), atol=1e-4)


def test_block_mask_hint_gives_identical_chains():
    """TXT->IMG noisy mean-field at a large batch: the clamp_suffix promise (no work on the clamped label block) must
    not change the result -- the random field is counter-addressed, so skipping the draws of clamped units shifts
    nothing."""
    import multimodal_idbn_b200 as M
    M.set_precision("tf32")
    try:
        V, H, Dz, K, B = 532, 256, 500, 32, 1024
        torch.manual_seed(0)
        r = M.RBM(V, H, 0.04, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(DEV)
        with torch.no_grad():
            r.W.data.mul_(4.0)
        y = torch.nn.functional.one_hot(torch.randint(0, K, (B,)), K).float().to(DEV)
        vk = torch.zeros(B, V, device=DEV); km = torch.zeros_like(vk); vk[:, Dz:] = y; km[:, Dz:] = 1
        r._mu_pull = {"mu_k": torch.rand(B, Dz, device=DEV), "eta0": 0.15}
        outs = []
        for hint in (-1, Dz):
            r.set_rng(77, 0)
            outs.append(r.noisy_meanfield_annealed(vk, km, n_steps=7, clamp_suffix=hint))
            r.set_rng(77, 0)
            outs.append(r.noisy_meanfield_annealed(vk, km, n_steps=1, T0=0.9, T1=0.9, sigma0=0.0, sharpen_last=0,
                                                   clamp_suffix=hint))
        # with the promise the chain runs in the persistent tensor-core kernel (chain_tc.cuh), without it step by step
        # on the stream passes: same tf32 operands, different rounding of the element-wise part
        torch.testing.assert_close(outs[2], outs[0], rtol=0, atol=2e-3)
        torch.testing.assert_close(outs[3], outs[1], rtol=0, atol=2e-3)
        assert torch.equal(outs[2][:, Dz:], y)
    finally:
        M.set_precision("fp32")


def test_c2_full_size_pipelined_matches_unpipelined(tmp_path, monkeypatch):
    """BASELINE config C2 at full size (iDBN [10000,1500,500], batch 64, tf32): the layer-pipelined step on SM
    partitions and the plain step give the same parameters up to the split-K summation order, and the losses
    written into pinned host memory equal the returned device losses."""
    monkeypatch.chdir(tmp_path)
    import multimodal_idbn_b200 as M
    M.set_precision("tf32")
    try:
        p = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95, LEARNING_RATE_DYNAMIC=True)
        xs = [(torch.rand(64, 10000, generator=torch.Generator().manual_seed(s)) < 0.1).float().to(DEV) for s in range(5)]
        runs = []
        for piped in (False, True):
            torch.manual_seed(0)
            m = M.iDBN([10000, 1500, 500], dict(p), None, None, torch.device(DEV))
            for i, l in enumerate(m.layers):
                l.set_rng(40 + i, 0)
            m.pipeline_layers = piped
            host = torch.full((4, 2), float("nan")).pin_memory()
            dev_losses = []
            for t in range(4):
                if t % 2 == 0:
                    m.train_step(xs[t], 0, 1, next_v=xs[t + 1], loss_out=host[t])
                else:
                    step = m.train_step(xs[t], 0, 1, next_v=xs[t + 1])
                    m.sync()
                    dev_losses.append(torch.stack(step).cpu())
            m.sync(); torch.cuda.synchronize()
            assert torch.isfinite(host[0]).all() and torch.isfinite(host[2]).all()
            runs.append((host[[0, 2]].clone(), torch.stack(dev_losses), [l.W.detach().cpu() for l in m.layers],
                         m.represent(xs[0]).cpu()))
        a, b = runs
        torch.testing.assert_close(b[0], a[0], rtol=2e-3, atol=1e-5)
        torch.testing.assert_close(b[1], a[1], rtol=2e-3, atol=1e-5)
        for wa, wb in zip(a[2], b[2]):
            # a hidden unit within rounding of its threshold may flip with the summation order: almost all entries tight
            assert float(((wa - wb).abs() > 1e-4).float().mean()) < 5e-3
        assert float((a[3] - b[3]).abs().mean()) < 2e-3
    finally:
        M.set_precision("fp32")


@pytest.mark.parametrize("B,n,pull", [(600, 20, True), (1031, 7, False), (2048, 50, True)])
def test_persistent_txt2img_chain_kernel_vs_oracle(M, B, n, pull):
    """k_chain_tc (chain_tc.cuh): TXT->IMG noisy mean-field annealing of many chains in ONE persistent tcgen05 kernel,
    chain state in shared memory, against the oracle fed the same random field (tf32 tolerance)."""
    V, H, Dz, K = 532, 256, 500, 32
    st, r = make(M, V, H, seed=23, scale=2.0, groups=[(Dz, V)])
    y = O.synthetic_labels(B, K, seed=3)
    mu = torch.rand(B, Dz, generator=torch.Generator().manual_seed(4))
    vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, Dz:] = y; km[:, Dz:] = 1
    ref = O.noisy_meanfield(st, vk, km, n_steps=n, mu_pull=(mu, 0.15) if pull else None, fld=RandomField(31, 0))
    r._mu_pull = {"mu_k": mu.to(DEV), "eta0": 0.15} if pull else None
    r.set_rng(31, 0)
    l0 = M.total_launches()
    out = r.noisy_meanfield_annealed(vk.to(DEV), km.to(DEV), n_steps=n, clamp_suffix=Dz)
    torch.cuda.synchronize()
    assert M.total_launches() - l0 == 1          # ONE kernel for the whole chain
    r._mu_pull = None
    torch.testing.assert_close(out.cpu(), ref, rtol=0, atol=3e-3)
    assert torch.equal(out.cpu()[:, Dz:], y)
