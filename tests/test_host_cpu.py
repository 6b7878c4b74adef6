"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
host logic of the drop-in classes (schedules, pickling, aliases, loud failure without CUDA), and the
data-parallel statistics exchange over a 2-process gloo group."""
import ctypes
import os
import pickle
import re
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(REPO, "include", "imdbn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imdbn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from multimodal_idbn_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/imdbn_b200.h but not exported"
    lib.imdbn_abi_version.restype = ctypes.c_int
    assert lib.imdbn_abi_version() == 1
    from multimodal_idbn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table covers the whole header


def test_sass_is_sm100a():
    from multimodal_idbn_b200.build import build
    out = subprocess.run(["cuobjdump", "-lelf", build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_ctx_create_without_gpu_fails_cleanly():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from multimodal_idbn_b200 import _lib
    lib = _lib.load_library()
    h = ctypes.c_void_p()
    assert lib.imdbn_ctx_create(ctypes.byref(h), 0) != 0 and not h.value


def test_no_cpu_fallback_and_reference_api_surface(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    import multimodal_idbn_b200 as M
    r = M.RBM(12, 5, 0.1, 1e-4, 0.5, softmax_groups=[(8, 12)]).to("cpu")
    for name in ("forward", "_visible_logits", "visible_probs", "sample_visible", "backward",
                 "backward_sample", "gibbs_step", "train_epoch", "_lin_schedule", "_hot_steps",
                 "conditional_gibbs_annealed", "noisy_meanfield_annealed", "conditional_gibbs",
                 "train_epoch_clamped"):
        assert callable(getattr(r, name))
    assert list(r.state_dict().keys()) == ["W", "hid_bias", "vis_bias"]
    assert not hasattr(r, "free_energy")                     # the reference ships without the hook
    with pytest.raises(RuntimeError, match="CUDA only"):
        r.forward(torch.zeros(2, 12))
    with pytest.raises(RuntimeError, match="CUDA only"):
        r.noisy_meanfield_annealed(torch.zeros(2, 12), torch.zeros(2, 12), n_steps=3)
    assert r._lin_schedule(0, 50, 3.0, 1.0) == 3.0 and r._lin_schedule(49, 50, 3.0, 1.0) == 1.0
    assert r._lin_schedule(3, 1, 3.0, 1.0) == 1.0 and r._hot_steps(50, 0.7) == 35
    p = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95,
             LEARNING_RATE_DYNAMIC=True, SPARSITY=True)
    d = M.iDBN([20, 10, 6], p, None, None, torch.device("cpu"))
    assert [l.sparsity for l in d.layers] == [False, True] and d.arch_str == "20-10-6"
    m = M.iMDBN([20, 10, 6], 4, params=p, device=torch.device("cpu"), num_labels=3)
    assert m.joint_rbm.num_visible == 9 and m.joint_rbm.softmax_groups == [(6, 9)]
    m2 = M.iMDBN([20, 10, 6], [5, 5], 4, params=p, device=torch.device("cpu"), num_labels=3,
                 logging_cfg={"x": 1})
    assert m2.joint_rbm.num_hidden == 4
    with pytest.raises(ValueError):
        M.iMDBN([20, 10], [5, 5], params=p)
    from imdbn.models.imdbn_bimodal import iMDBN_BiModal
    b = iMDBN_BiModal([36, 20, 12], [28, 16, 10], 14, params=p, device=torch.device("cpu"))
    assert b is not None and iMDBN_BiModal is M.iMDBN_BiModal and b.num_joint_layers == 1
    assert b.joint_rbm.num_visible == 22 and b.joint_rbm.softmax_groups == [] and b.arch_str.endswith("JOINT14")
    for name in ("load_pretrained_mod1_dbn", "load_pretrained_mod2_dbn", "init_joint_bias_from_data",
                 "_cross_reconstruct", "represent", "train_joint", "save_model", "load_model"):
        assert callable(getattr(b, name))


def test_pickle_layout_and_aliases(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    import multimodal_idbn_b200 as M
    import imdbn.models as IM
    import imdbn.models.gdbn_model_complete as MONO
    assert IM.RBM is M.RBM and MONO.iMDBN is M.iMDBN and M.RBM.__module__ == "imdbn.models.rbm"
    assert sys.modules["src.classes.rbm_model"].RBM is M.RBM     # legacy Groundeep aliases
    p = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95,
             LEARNING_RATE_DYNAMIC=True)
    m = M.iMDBN([20, 10, 6], 4, params=p, device=torch.device("cpu"), num_labels=3)
    m.z_class_mean = torch.zeros(3, 6)
    m.save_model(str(tmp_path / "m.pkl"))
    blob = open(tmp_path / "m.pkl", "rb").read()
    assert b"imdbn.models.rbm" in blob and b"multimodal_idbn_b200" not in blob
    d = M.iMDBN.load_model(str(tmp_path / "m.pkl"), device=torch.device("cpu"))
    assert set(d) >= {"layers", "params", "image_idbn", "joint_rbm", "num_labels", "Dz_img",
                      "arch_str", "features", "metadata", "z_class_mean"}
    r = d["joint_rbm"]
    for k in ("W", "hid_bias", "vis_bias"):
        assert isinstance(getattr(r, k), torch.nn.Parameter)
    for k in ("W_m", "hb_m", "vb_m", "num_visible", "num_hidden", "lr", "weight_decay", "momentum",
              "dynamic_lr", "final_momentum", "sparsity", "sparsity_factor", "softmax_groups"):
        assert k in r.__dict__
    assert "_stats_buf" not in r.__dict__
    # a {"layers": ...} checkpoint feeds load_pretrained_image_idbn
    m.image_idbn.save_model(str(tmp_path / "i.pkl"))
    assert m.load_pretrained_image_idbn(str(tmp_path / "i.pkl"))
    assert not m.load_pretrained_image_idbn(str(tmp_path / "missing.pkl"))


WORKER = r"""
import os, sys, torch, torch.distributed as td
sys.path.insert(0, {repo!r})
from oracle import rbm_oracle as O
from oracle.philox import RandomField
import multimodal_idbn_b200.dist as D
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
td.init_process_group("gloo", rank=rank, world_size=world)
D.enable()
assert D.state().rank == rank and D.state().world == world
torch.set_num_threads(1)
V, H, B = 40, 24, 12
st = O.new_state(V, H, seed=3, groups=[(32, 40)], lr=0.1, weight_decay=1e-4, momentum=0.5,
                 final_momentum=0.95, dynamic_lr=True, sparsity=True, sparsity_factor=0.1)
st.W *= 3
data = torch.cat([O.synthetic_images(B, 32, p=0.4, seed=5), O.synthetic_labels(B, 8, seed=6)], 1)
full = st.clone()
loss_full, _ = O.cd_train(full, data, 2, 2, RandomField(9, 4))
lo, hi = D.shard_rows(B, rank, world)
s = O.cd_statistics(st, data[lo:hi], 2, RandomField(9, 4), row0=lo)       # global-row addressing
flat = torch.cat([s["dS"].reshape(-1), s["dh"], s["dv"], s["pos_h_sum"], s["sq_err"].reshape(1)])
D.state().all_reduce(flat)
n = V * H
lr, mom = O.lr_and_momentum(st, 2)
O.apply_update(st, flat[:n].view(V, H), flat[n:n + H], flat[n + H:n + H + V], B, lr, mom,
               pos_h_mean=flat[n + H + V:n + 2 * H + V] / B)
loss = flat[-1] / (B * V)
for a, b in ((st.W, full.W), (st.hb, full.hb), (st.vb, full.vb), (st.Wm, full.Wm), (loss, loss_full)):
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
td.destroy_process_group()
print("rank", rank, "ok")
"""


def test_data_parallel_statistics_two_ranks_gloo(tmp_path):
    """World size 2 over gloo: shard the minibatch, all-reduce the CD statistics through
    ``multimodal_idbn_b200.dist``, apply the update -> identical to the single-process update."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), LOCAL_RANK=str(r), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out
        assert "ok" in out


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_pickle_resolves_to_drop_in_classes():
    """tests/golden/ref_idbn.pkl was written by the reference's iDBN.save_model (reference objects)."""
    import multimodal_idbn_b200 as M
    with open(os.path.join(REPO, "tests", "golden", "ref_idbn.pkl"), "rb") as f:
        d = pickle.load(f)
    assert set(d) == {"layers", "params"} and len(d["layers"]) == 2
    r = d["layers"][0]
    assert type(r) is M.RBM and r.num_visible == 40 and r.num_hidden == 20
    assert isinstance(r.W, torch.nn.Parameter) and r.W_m.shape == r.W.shape
    assert not hasattr(r, "_rng_seed")                 # a reference object: no drop-in extras yet
    r._next_rng()                                      # ...which the drop-in creates on first use
    assert r._rng_stream == 1


def test_device_resident_dataset_and_loaders(tmp_path):
    """datasets.py host logic, on the CPU device: npz loading, one-hot labels from numerosities,
    seeded split, Subset-like validation set with the feature lists iDBN.__init__ reads (idbn.py:131-137), one
    permutation per epoch, contiguous batch views with a ragged last batch, and the reference scripts' import path."""
    import numpy as np
    import torch
    from imdbn.datasets.uniform_dataset import create_dataloaders_uniform
    from multimodal_idbn_b200.datasets import DeviceLoader, DeviceDataset
    rng = np.random.default_rng(0)
    N, K = 53, 8
    imgs = (rng.random((N, 10, 10)) < 0.1).astype(np.uint8)
    nums = rng.integers(1, K + 1, size=N)
    np.savez(tmp_path / "toy.npz", D=imgs, N_list=nums, cumArea_list=rng.random(N), CH_list=rng.random(N),
             density_list=rng.random(N))
    tr, va, te = create_dataloaders_uniform(str(tmp_path), "toy.npz", batch_size=8, num_workers=3, multimodal_flag=True,
                                            device="cpu", seed=3)
    assert len(tr.dataset) + len(va.dataset) + len(te.dataset) == N
    assert sorted(tr.dataset.indices + va.dataset.indices + te.dataset.indices) == list(range(N))
    base = va.dataset.dataset
    assert len(base.labels) == N and len(base.cumArea_list) == N and len(base.CH_list) == N and len(base.density_list) == N
    assert [base.labels[i] for i in va.dataset.indices] == [int(nums[i]) for i in va.dataset.indices]
    # epoch = one permutation; every sample exactly once; last batch ragged; labels follow their images
    seen, batches = [], list(tr)
    assert len(batches) == len(tr) and batches[-1][0].shape[0] == len(tr.dataset) - 8 * (len(tr) - 1)
    for xb, yb in batches:
        assert xb.shape[1:] == (10, 10) and yb.shape[1] == K and xb.dtype == torch.float32 and xb.is_contiguous()
        for x, y in zip(xb, yb):
            hits = [i for i in tr.dataset.indices if np.array_equal(imgs[i], x.numpy().astype(np.uint8))
                    and int(nums[i]) - 1 == int(y.argmax())]
            assert hits
            seen.append(hits[0])
    first = torch.cat([b[0] for b in batches])
    second = torch.cat([b[0] for b in tr])
    assert first.shape == second.shape and not torch.equal(first, second)          # reshuffled next epoch
    assert torch.equal(torch.cat([b[0] for b in va]), torch.cat([b[0] for b in va]))  # validation order is fixed
    # a plain loader over a whole dataset, flat views
    ds = DeviceDataset(torch.from_numpy(imgs), torch.from_numpy(nums), "cpu")
    xb, yb = next(iter(DeviceLoader(ds, 5, flat=True)))
    assert xb.shape == (5, 100) and torch.equal(xb, ds.images[:5]) and xb.data_ptr() == ds.images.data_ptr()


def test_probe_utils_binning_split_and_probe_match_the_reference():
    """quantile binning, bin names, the stratified split and the early-stopped linear probe of
    probe_utils.py reproduce what the unmodified reference returned on the same seeded inputs (tests/golden/probe.npz,
    written by make_golden.case_probe).  Everything here runs on the CPU device; the GPU test covers the device path."""
    import numpy as np
    from multimodal_idbn_b200 import probe_utils as P
    import imdbn.utils.probe_utils as alias
    assert alias.log_linear_probe is P.log_linear_probe and alias.stratified_split is P.stratified_split
    g = np.load(os.path.join(REPO, "tests", "golden", "probe.npz"))
    for key in ("cont", "ties", "labels"):
        v = torch.from_numpy(g[key + "_values"])
        y, edges = P.make_bin_labels(v, n_bins=5)
        assert torch.equal(y, torch.from_numpy(g[key + "_bins"]))
        assert torch.equal(edges, torch.from_numpy(g[key + "_edges"]))
        tr, te = P.stratified_split(y, test_size=0.2, rng_seed=42)
        assert tr == g[key + "_train_idx"].tolist() and te == g[key + "_test_idx"].tolist()
        assert P._format_bin_names(edges, precision=4) == g[key + "_names"].tolist()
    X, y = torch.from_numpy(g["probe_X"]), torch.from_numpy(g["probe_y"])
    tr, te = P.stratified_split(y, test_size=0.2, rng_seed=42)
    assert tr == g["probe_train_idx"].tolist() and te == g["probe_test_idx"].tolist()
    torch.manual_seed(int(g["probe_seed"]))
    acc, y_true, y_pred = P.train_linear_classifier(X[tr], y[tr], X[te], y[te], device=torch.device("cpu"), n_classes=5,
                                                    max_steps=300, lr=1e-2, weight_decay=0.0, patience=20, min_delta=0.0)
    assert y_true == g["probe_y_true"].tolist()
    assert y_pred == g["probe_y_pred"].tolist()            # same stopping step, same selected parameters
    assert abs(acc - float(g["probe_acc"])) < 1e-7
    cm = P.confusion_matrix(y_true, y_pred, 5)
    assert int(cm.sum()) == len(y_true) and int(cm.diag().sum()) == round(acc * len(y_true))


def test_probe_pca_projection_matches_sklearn():
    """probe_utils.pca_project (the device replacement of the PCA(n).fit_transform calls of idbn.py:262-283) against
    scikit-learn on the same matrix: projections agree up to the sign of each component, variance ratios exactly."""
    import numpy as np
    from sklearn.decomposition import PCA
    from multimodal_idbn_b200.probe_utils import pca_project
    g = torch.Generator().manual_seed(3)
    X = torch.randn(200, 24, generator=g) @ torch.diag(torch.linspace(3.0, 0.1, 24)) + 0.5
    for k in (2, 3):
        Z, ratio = pca_project(X, k)
        ref = PCA(n_components=k).fit(X.numpy())
        assert np.allclose(np.abs(Z.numpy()), np.abs(ref.transform(X.numpy())), atol=2e-4)
        assert np.allclose(ratio.numpy(), ref.explained_variance_ratio_, atol=1e-5)
