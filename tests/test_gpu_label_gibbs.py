"""IMG->TXT label-only conditional-Gibbs kernel (label_gibbs.cuh) against the oracle and against the
generic persistent kernel."""
import pytest
import torch

from oracle import rbm_oracle as O
from oracle.philox import RandomField

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("prec,tol", [("fp32", 3e-5), ("tf32", 3e-3), ("tf32x2", 3e-5)])
@pytest.mark.parametrize("B,K,n", [(96, 32, 50), (7, 32, 0), (300, 20, 9)])
def test_label_only_matches_oracle_and_generic_path(prec, tol, B, K, n):
    import multimodal_idbn_b200 as M
    M.set_precision(prec)
    try:
        Dz, H = 500, 256
        V = Dz + K
        st = O.new_state(V, H, seed=5, groups=[(Dz, V)])
        st.W *= 2.0
        g = torch.Generator().manual_seed(1)
        st.hb.copy_(torch.randn(H, generator=g) * 0.1); st.vb.copy_(torch.randn(V, generator=g) * 0.1)
        r = M.RBM(V, H, 0.1, 1e-4, 0.5, softmax_groups=[(Dz, V)]).to(DEV)
        with torch.no_grad():
            r.W.data.copy_(st.W); r.hid_bias.data.copy_(st.hb); r.vis_bias.data.copy_(st.vb)
        z = torch.rand(B, Dz, generator=g)
        vk = torch.zeros(B, V); km = torch.zeros(B, V); vk[:, :Dz] = z; km[:, :Dz] = 1
        ref = O.conditional_gibbs(st, vk, km, n_steps=n, fld=RandomField(11, 3))
        r.set_rng(11, 3)
        fast = r.conditional_gibbs(vk.to(DEV), km.to(DEV), n_steps=n, clamp_prefix=Dz)
        r.set_rng(11, 3)
        slow = r.conditional_gibbs(vk.to(DEV), km.to(DEV), n_steps=n)
        torch.testing.assert_close(fast.cpu(), ref, rtol=0, atol=tol)
        torch.testing.assert_close(fast, slow, rtol=0, atol=tol)
        assert torch.allclose(fast[:, Dz:].sum(1).cpu(), torch.ones(B), atol=1e-4)
    finally:
        M.set_precision("fp32")
